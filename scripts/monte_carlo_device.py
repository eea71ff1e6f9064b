"""Monte-Carlo FER run with frames synthesized ON THE DEVICE: distinct frames (the reference harness'
streams, keyed by absolute frame index), no host feed.  README-shaped N = 2^20 stand-in code."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness

sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 0.824
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
log2p = int(sys.argv[3]) if len(sys.argv) > 3 else 12
gen = alist.make_readme_shaped_code(20, seed=1)
dec = L.Decoder(L.Code.from_generated(gen), L.FP16, log2p)
P = dec.parallel_factor()
n = 2 * P  # frames per batch
N, M = gen.n_vars, gen.n_checks
W, S = N // 32, (M + 31) // 32
d_in = torch.empty((N, n), dtype=torch.float16, device="cuda")
d_syn = torch.empty((n, S), dtype=torch.int32, device="cuda")
d_ref = torch.empty((n, W), dtype=torch.int32, device="cuda")
d_res = torch.empty((n, W), dtype=torch.int32, device="cuda")
sig_h = float(np.float16(sigma))  # the reference's fp16 build parses sigma into a half
prm = harness.channel_param("awgn", sig_h)
tot = {"frames": 0, "frame_errors": 0, "frame_errors_gt15": 0, "bit_errors": 0, "synth_s": 0.0, "decode_s": 0.0}
hist = {}
for b in range(batches):
    torch.cuda.synchronize(); t = time.perf_counter()
    dec.synthesize_frames(L.AWGN, sig_h, b * n, n, d_in, d_syn, d_ref)
    torch.cuda.synchronize(); tot["synth_s"] += time.perf_counter() - t; t = time.perf_counter()
    _, iters, rep = dec.decode(L.AWGN, prm, d_in, d_syn, n_frames=n, num_iter_max=120, results=d_res)
    torch.cuda.synchronize(); tot["decode_s"] += time.perf_counter() - t
    x = torch.bitwise_xor(d_res, d_ref).view(torch.uint8)
    err = torch.zeros(n, dtype=torch.int64, device="cuda")
    for k in range(8):
        err += ((x >> k) & 1).sum(dim=1, dtype=torch.int64)
    err = err.cpu().numpy()
    tot["frames"] += n; tot["frame_errors"] += int((err > 0).sum()); tot["frame_errors_gt15"] += int((err > 15).sum())
    tot["bit_errors"] += int(err.sum())
    for k, v in zip(*np.unique(iters, return_counts=True)):
        hist[int(k)] = hist.get(int(k), 0) + int(v)
    print(f"batch {b}: frames {n} avg_iter {rep['avg_iter']:.2f} frame_errors {(err>0).sum()} synth {tot['synth_s']:.2f}s decode {tot['decode_s']:.2f}s", flush=True)
tot.update({"sigma": sig_h, "frames_in_flight": P, "fer": tot["frame_errors"] / tot["frames"], "fer_gt15": tot["frame_errors_gt15"] / tot["frames"],
            "ber": tot["bit_errors"] / (tot["frames"] * N), "iter_hist": dict(sorted(hist.items())),
            "avg_iter": sum(k * v for k, v in hist.items()) / tot["frames"],
            "decode_mbps": tot["frames"] * N / 2**20 / tot["decode_s"], "synth_mbps": tot["frames"] * N / 2**20 / tot["synth_s"]})
print(json.dumps(tot))
json.dump(tot, open("gpurun_out/r01_monte_carlo_device.json", "w"), indent=1)
