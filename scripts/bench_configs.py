"""Extra bench lines for BASELINE.json configs[1] and configs[4] (the headline bench.py line is configs[2]).
Device-resident frames synthesized on the GPU (the reference harness' streams), reference scheduling semantics."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
out = []
CASES = [
    ("configs[1] shape: rate-0.9 code on a BSC, fp32, 2^10 in flight (stand-in: (3,32)-regular N=2^20, p=0.004, loading 4)",
     lambda: alist.make_regular_code(1 << 20, 3, 32, seed=4), L.FP32, 10, 4, "bsc", 0.004, 100, 128),
    ("configs[4] shape: regular (3,6) N=2^22, AWGN sigma=0.875 (threshold 0.881), up to 500 iterations, fp16, loading 4",
     lambda: alist.make_regular_code(1 << 22, 3, 6, seed=6), L.FP16, 11, 4, "awgn", 0.875, 500, 64),
    ("configs[4] shape, fp32",
     lambda: alist.make_regular_code(1 << 22, 3, 6, seed=6), L.FP32, 10, 4, "awgn", 0.875, 500, 64),
]
for name, make, prec, log2p, load, channel, noise, imax, pool in CASES:
    gen = make()
    dec = L.Decoder(L.Code.from_generated(gen), prec, log2p)
    P = dec.parallel_factor(); n = P * load
    N, M, E = gen.n_vars, gen.n_checks, gen.n_edges
    dt = torch.float16 if prec == L.FP16 else torch.float32
    d_in = torch.empty((N, pool), dtype=dt, device="cuda"); d_syn = torch.empty((pool, (M + 31) // 32), dtype=torch.int32, device="cuda")
    d_ref = torch.empty((pool, N // 32), dtype=torch.int32, device="cuda"); d_res = torch.empty((n, N // 32), dtype=torch.int32, device="cuda")
    ch = L.BSC if channel == "bsc" else L.AWGN
    nz = float(np.float16(noise)) if (prec == L.FP16 and channel == "awgn") else noise
    dec.synthesize_frames(ch, nz, 0, pool, d_in, d_syn, d_ref)
    prm = harness.channel_param(channel, nz)
    kw = dict(n_frames=n, num_iter_max=imax, results=d_res, pool_frames=pool)
    dec.decode(ch, prm, d_in, d_syn, **kw)  # warm-up
    torch.cuda.synchronize(); t = time.perf_counter()
    _, iters, rep = dec.decode(ch, prm, d_in, d_syn, flags=L.FLAG_TIME_KERNELS, **kw)
    torch.cuda.synchronize(); el = time.perf_counter() - t
    s = 2 if prec == L.FP16 else 4
    b_it = (4 * E + N) * s + 4 * ((M + 31) // 32)
    conv = iters < imax
    idx = np.arange(n) % pool
    ok = bool((d_res.cpu().numpy()[conv] == d_ref.cpu().numpy()[idx[conv]]).all())
    row = {"workload": name, "frames_in_flight": P, "frames": n, "value_mbps": n * N / 2**20 / el, "ms": 1e3 * el,
           "avg_iter": rep["avg_iter"], "max_iter": rep["max_iter"], "frac_converged": float(conv.mean()), "converged_frames_correct": ok,
           "roofline_frac": b_it * rep["slot_iterations"] / (rep["cn_s"] + rep["vn_s"]) / 1e9 / PEAK,
           "code": {"N": N, "M": M, "E": E}}
    out.append(row); print(json.dumps(row), flush=True)
    del dec, d_in, d_syn, d_ref, d_res; torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/" + (sys.argv[1] if len(sys.argv) > 1 else "r01") + "_bench_other_configs.json", "w"), indent=1)
