"""Generates tests/golden/*.npz by running the UNMODIFIED reference CUDA backend (oracle/_ref, built from
/root/reference by oracle/Makefile) on a B200:   gpurun -- python tests/golden/make_golden.py
Each fixture holds the alist text, the inputs, and the reference's outputs (messages after fixed iteration
counts, hard decisions, violation flags, decode() results and iteration statistics)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ldpc_decoder_b200 import alist, harness  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)

CASES = [
    # name, generator, channel, noise (kernel steps), noise (decode), frames for decode, log2P, imax, k
    ("reg36_n384", lambda: alist.make_regular_code(384, 3, 6, seed=21), "awgn", 0.80, 0.74, 24, 3, 40, 5),
    ("punct_n512", lambda: alist.make_readme_shaped_code(9, seed=22), "awgn", 0.70, 0.50, 20, 2, 60, 10),
    ("hideg_bsc_n256", lambda: alist.make_irregular_code(256, 32, [(224, 3), (16, 10), (16, 12)], 0, seed=23),
     "bsc", 0.02, 0.01, 12, 3, 30, 3),
]

for prec in ("fp32", "fp16"):
    dt = np.float32 if prec == "fp32" else np.float16
    for name, make, channel, noise_k, noise_d, n_dec, log2p, imax, k in CASES:
        gen = make()
        text = alist.alist_text(gen)
        d = tempfile.mkdtemp()
        path = os.path.join(d, "c.alist")
        open(path, "w").write(text)
        n_reg = gen.n_vars - gen.n_erased
        P = 1 << log2p
        out = {"alist": np.frombuffer(text.encode(), dtype=np.uint8), "channel": channel, "log2p": log2p,
               "imax": imax, "k": k, "noise_k": noise_k, "noise_d": noise_d}
        # kernel-level trace
        inp, syn, bits = harness.make_frames(gen, P, channel, noise_k, seed=31, dtype=dt)
        ref = rb.RefDecoder(path, channel, noise_k, log2p, prec)
        out["param_k"] = np.float32(ref.channel_param())
        out["k_inputs"], out["k_syndromes"] = inp, syn
        ref.refill(inp[:n_reg], syn)
        out["msg_it0"] = np.stack([ref.messages(s) for s in range(P)])
        out["priors"] = np.stack([ref.priors(s) for s in range(P)])
        done = 0
        for target in (1, 2, 5, 10):
            for _ in range(target - done):
                ref.backward()
                ref.forward(True)
            done = target
            out[f"msg_it{target}"] = np.stack([ref.messages(s) for s in range(P)])
            out[f"hard_it{target}"] = ref.deinterlace()
            out[f"viol_it{target}"] = ref.check()
        ref.close()
        # end-to-end decode()
        inp, syn, bits = harness.make_frames(gen, n_dec, channel, noise_d, seed=32, dtype=dt)
        ref = rb.RefDecoder(path, channel, noise_d, log2p, prec)
        out["param_d"] = np.float32(ref.channel_param())
        res, rep = ref.decode(inp, syn, imax, k)
        ref.close()
        out["d_inputs"], out["d_syndromes"], out["d_bits"], out["d_results"] = inp, syn, bits, res
        out["d_report"] = np.array([rep["max_iter"], rep["min_iter"], rep["avg_iter"]], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, f"{name}_{prec}.npz"), **out)
        print(name, prec, "decode report", rep, "errors", harness.count_bit_errors(res, bits).sum())
