"""Generates tests/golden/harness_*.npz from the UNMODIFIED reference test-data generator pieces (ChaCha8 PRNG,
rng.h, channel.cpp, bool_vec.h, ldpc_code.cpp) compiled into oracle/_ref/libldpc_ref_harness.so (CPU only):
    make -C oracle refharness && python tests/golden/make_golden_harness.py"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ldpc_decoder_b200 import alist  # noqa: E402

L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libldpc_ref_harness.so"))
L.refh_capacity.restype = C.c_float
L.refh_capacity.argtypes = [C.c_int, C.c_float]
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731

out = {}
for seed in (0, 1, 511, (1 << 32) | 7, 0xDEADBEEFCAFE):
    n = 2000
    a = np.zeros(n, np.uint32); L.refh_random_ints(C.c_uint64(seed), n, P(a))
    u = np.zeros(n, np.float32); L.refh_units(C.c_uint64(seed), n, P(u))
    g = np.zeros(n, np.float32); L.refh_gaussians(C.c_uint64(seed), n, P(g))
    out[f"ints_{seed}"], out[f"units_{seed}"], out[f"gauss_{seed}"] = a, u, g
out["capacity"] = np.array([[k, x, L.refh_capacity(k, C.c_float(x))] for k, x in
                            ((0, 0.085), (0, 0.01), (1, 0.94), (1, 0.5), (1, 0.0005))], dtype=np.float64)
gen = alist.make_readme_shaped_code(9, seed=22)
text = alist.alist_text(gen)
d = tempfile.mkdtemp(); path = os.path.join(d, "c.alist"); open(path, "w").write(text)
out["alist"] = np.frombuffer(text.encode(), dtype=np.uint8)
for kind, noise, start, n_vec, batch in ((1, 0.94, 0, 40, 0), (1, 0.7, 100, 33, 2), (0, 0.085, 5, 64, 1)):
    N, M = gen.n_vars, gen.n_checks
    noisy = np.zeros((N, n_vec), np.float32)
    ref = np.zeros((n_vec, N // 32), np.uint32)
    syn = np.zeros((n_vec, (M + 31) // 32), np.uint32)
    assert L.refh_create_data(path.encode(), kind, C.c_float(noise), start, n_vec, batch, P(noisy), P(ref), P(syn)) == 0
    tag = f"data_{kind}_{start}_{n_vec}_{batch}"
    out[tag + "_noise"] = np.float32(noise)
    out[tag + "_noisy"], out[tag + "_ref"], out[tag + "_syn"] = noisy, ref, syn
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "harness_reference.npz"), **out)
print("written", len(out), "arrays")
