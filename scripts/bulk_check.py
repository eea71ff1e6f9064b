"""Experiment: bulk-async (TMA) staged check-node kernel vs the register kernel: bit-exact outputs + speed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness

gen = alist.make_readme_shaped_code(16, seed=1)
code = L.Code.from_generated(gen)
for prec, dt, log2p in ((L.FP16, np.float16, 11), (L.FP32, np.float32, 10)):
    P = 1 << log2p
    inp, syn, bits = harness.make_frames(gen, 64, "awgn", 0.8, seed=1, dtype=dt)
    n_reg = gen.n_vars - gen.n_erased
    st = np.tile(inp[:n_reg], (1, P // 64)); sy = np.tile(syn, (P // 64, 1))
    os.environ.pop("LDPC_B200_CN_BULK", None)
    a = L.Decoder(code, prec, log2p)
    os.environ["LDPC_B200_CN_BULK"] = "1"
    b = L.Decoder(code, prec, log2p)
    for d in (a, b):
        d.refill(L.AWGN, harness.channel_param("awgn", 0.8), st, sy)
        d.iterate(7, final_bits=True)
    same = all(np.array_equal(a.messages(s).view(np.uint32), b.messages(s).view(np.uint32)) for s in (0, 5, 63, 64, P - 1))
    print("fp16" if prec == L.FP16 else "fp32", "bulk == register kernel:", same, flush=True)
