"""Same GPU, same inputs: the UNMODIFIED reference CUDA backend (oracle/_ref, sm_100 build) vs this decoder.
README sample-run shape (-p 8 -m 2 -i 120) and the largest size the reference can index (P*E < 2^32)."""
import os, sys, time, tempfile, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness
from oracle import ref_binding as rb

gen = alist.make_readme_shaped_code(20, seed=1)
d = tempfile.mkdtemp(); path = os.path.join(d, "c.alist"); alist.write_alist(gen, path)
sigma = 0.824
out = []
for prec, log2p, load in (("fp16", 8, 2), ("fp32", 8, 2), ("fp16", 10, 2), ("fp32", 10, 2)):
    dt = np.float16 if prec == "fp16" else np.float32
    n = (1 << log2p) * load
    pool = 128
    inp, syn, bits = harness.make_frames(gen, pool, "awgn", sigma, seed=3, dtype=dt)
    reps = n // pool
    inp_full = np.ascontiguousarray(np.tile(inp, (1, reps))); syn_full = np.ascontiguousarray(np.tile(syn, (reps, 1)))
    prm = harness.channel_param("awgn", sigma)
    ref = rb.RefDecoder(path, "awgn", sigma, log2p, prec)
    ref.decode(inp_full, syn_full, 120, 10)  # warm-up
    t = time.perf_counter(); rres, rrep = ref.decode(inp_full, syn_full, 120, 10); t_ref = time.perf_counter() - t
    ref.close()
    dec = L.Decoder(L.Code(path=path), L.FP16 if prec == "fp16" else L.FP32, log2p)
    dec.decode(L.AWGN, prm, inp_full, syn_full, num_iter_max=120)
    t = time.perf_counter(); res, iters, rep = dec.decode(L.AWGN, prm, inp_full, syn_full, num_iter_max=120); t_new = time.perf_counter() - t
    same = bool(np.array_equal(res, rres))
    row = {"precision": prec, "frames_in_flight": 1 << log2p, "frames": n, "avg_iter": rep["avg_iter"],
           "reference_build_mbps": n * gen.n_vars / 2**20 / t_ref, "b200_native_mbps": n * gen.n_vars / 2**20 / t_new,
           "speedup": t_ref / t_new, "decisions_identical": same, "ref_avg_iter": rrep["avg_iter"]}
    out.append(row); print(json.dumps(row), flush=True)
    del dec
json.dump(out, open("gpurun_out/r06_vs_reference_build.json", "w"), indent=1)
