"""CPU tests of the CLI harness' data generation against fixtures produced by the reference's own
PRNG / channel / bit-vector code (tests/golden/make_golden_harness.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "harness_reference.npz")
HARNESS_DIR = os.path.join(ROOT, "ldpc_decoder_b200", "csrc", "harness")


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-s", "-C", HARNESS_DIR])
    L = C.CDLL(os.path.join(ROOT, "ldpc_decoder_b200", "libldpc_b200_harness.so"))
    L.ldpc_b200_harness_capacity.restype = C.c_float
    L.ldpc_b200_harness_capacity.argtypes = [C.c_int, C.c_float]
    L.ldpc_b200_harness_float_to_half.restype = C.c_uint16
    L.ldpc_b200_harness_float_to_half.argtypes = [C.c_float]
    L.ldpc_b200_harness_half_to_float.restype = C.c_float
    L.ldpc_b200_harness_half_to_float.argtypes = [C.c_uint16]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_chacha8_prng_streams_match_reference(lib):
    z = np.load(GOLD)
    seeds = sorted({int(k.split("_")[1]) for k in z.files if k.startswith("ints_")})
    assert len(seeds) >= 5
    for seed in seeds:
        n = z[f"ints_{seed}"].size
        a = np.zeros(n, np.uint32); lib.ldpc_b200_harness_random_ints(C.c_uint64(seed), n, _p(a))
        u = np.zeros(n, np.float32); lib.ldpc_b200_harness_units(C.c_uint64(seed), n, _p(u))
        g = np.zeros(n, np.float32); lib.ldpc_b200_harness_gaussians(C.c_uint64(seed), n, _p(g))
        assert np.array_equal(a, z[f"ints_{seed}"]), seed       # crosses a 1536-byte refill (384 words)
        assert np.array_equal(u, z[f"units_{seed}"]), seed
        assert np.array_equal(g.view(np.uint32), z[f"gauss_{seed}"].view(np.uint32)), seed


def test_capacity_matches_reference(lib):
    for kind, x, cap in np.load(GOLD)["capacity"]:
        assert lib.ldpc_b200_harness_capacity(int(kind), C.c_float(x)) == pytest.approx(cap, rel=1e-6)


def test_create_data_matches_reference(lib, tmp_path):
    z = np.load(GOLD)
    path = str(tmp_path / "c.alist")
    open(path, "wb").write(bytes(z["alist"]))
    tags = sorted({k[:-6] for k in z.files if k.endswith("_noisy")})
    assert len(tags) == 3
    for tag in tags:
        _, kind, start, n_vec, batch = tag.split("_")
        kind, start, n_vec, batch = int(kind), int(start), int(n_vec), int(batch)
        noisy = np.zeros_like(z[tag + "_noisy"]); ref = np.zeros_like(z[tag + "_ref"]); syn = np.zeros_like(z[tag + "_syn"])
        rc = lib.ldpc_b200_harness_create_data(path.encode(), kind, C.c_float(float(z[tag + "_noise"])), start, n_vec,
                                               batch, 0, _p(noisy), _p(ref), _p(syn))
        assert rc == 0
        assert np.array_equal(ref, z[tag + "_ref"]) and np.array_equal(syn, z[tag + "_syn"]), tag
        assert np.array_equal(noisy.view(np.uint32), z[tag + "_noisy"].view(np.uint32)), tag


def test_half_conversion_matches_numpy(lib):
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1e-7, 1e-4, 1, 100, 70000)])
    vals = np.concatenate([vals, np.array([0.0, -0.0, 65504.0, 65520.0, 65519.9, 5.96e-8, 2.98e-8, 2.9802322e-8,
                                           6.1035156e-5, 6.0975552e-5], np.float32)])
    for v in vals[::37]:
        assert lib.ldpc_b200_harness_float_to_half(C.c_float(float(v))) == int(np.float32(v).astype(np.float16).view(np.uint16))
    for h in range(0, 0x7C00, 97):
        assert lib.ldpc_b200_harness_half_to_float(h) == float(np.uint16(h).view(np.float16))
