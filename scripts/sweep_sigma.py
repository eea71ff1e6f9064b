"""Find the noise level at which a stand-in code behaves like the README run (avg ~90 iterations, small FER)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sigmas = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.6, 0.7, 0.75, 0.8, 0.85]
kind = sys.argv[3] if len(sys.argv) > 3 else "readme"
n_frames = 64
if kind == "readme":
    gen = alist.make_readme_shaped_code(log2n, seed=1)
else:
    gen = alist.make_regular_code(1 << log2n, 3, 6, seed=1)
print("code", gen.n_vars, gen.n_checks, gen.n_edges, "erased", gen.n_erased, flush=True)
code = L.Code.from_generated(gen)
dec = L.Decoder(code, L.FP16, 6)
np.save("gpurun_out/phi_table_fp16.npy", dec.phi_table_fp16())
for s in sigmas:
    inp, syn, bits = harness.make_frames(gen, n_frames, "awgn", s, seed=1, dtype=np.float16)
    t = time.time()
    res, iters, rep = dec.decode(L.AWGN, harness.channel_param("awgn", s), inp, syn, num_iter_max=300, num_iter_check_parity=10)
    err = harness.count_bit_errors(res, bits)
    print(f"sigma {s}: avg_iter {rep['avg_iter']:.1f} min {rep['min_iter']} max {rep['max_iter']} frames_with_err {(err>0).sum()}/{n_frames} "
          f"max_err {err.max()} device_s {rep['device_s']:.3f} ({time.time()-t:.2f}s wall)", flush=True)
