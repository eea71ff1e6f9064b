import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def small_codes(tmp_path_factory):
    """name -> (generated Code, alist path).  Shapes chosen to hit every kernel path:
    regular (3,6); punctured irregular (README shape, degrees <= 6); high-degree rows (> 8, the
    re-read path); tiny code where one warp spans several rows."""
    from ldpc_decoder_b200 import alist

    d = tmp_path_factory.mktemp("codes")
    codes = {
        "reg36": alist.make_regular_code(2048, 3, 6, seed=3),
        "readme13": alist.make_readme_shaped_code(13, seed=5),
        "hideg": alist.make_irregular_code(1024, 128, [(896, 3), (64, 10), (64, 12)], n_erased=0, seed=7),
        "tiny": alist.make_regular_code(96, 3, 6, seed=11),
    }
    out = {}
    for name, c in codes.items():
        p = str(d / f"{name}.alist")
        alist.write_alist(c, p)
        out[name] = (c, p)
    return out
