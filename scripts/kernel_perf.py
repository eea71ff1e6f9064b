"""Quick per-pass bandwidth of the CN / VN kernels through the kernel-level C-ABI entry points."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness

configs = [a for a in sys.argv[1:] if not a.startswith(("code=", "order="))] or ["fp16:10", "fp16:12", "fp32:10"]
order_arg = ([a[6:] for a in sys.argv[1:] if a.startswith("order=")] or ["natural"])[0]  # natural | locality (in-decoder)
code_arg = ([a[5:] for a in sys.argv[1:] if a.startswith("code=")] or ["readme:20"])[0].split(":")
if code_arg[0] == "readme":
    gen = alist.make_readme_shaped_code(int(code_arg[1]), seed=1)
elif code_arg[0] in ("banded", "scrambled"):  # banded:log2n:dv:dc:window  (locality experiment: edges stay within a window)
    n, dv, dc, win = 1 << int(code_arg[1]), int(code_arg[2]), int(code_arg[3]), int(code_arg[4])
    rng = np.random.default_rng(1)
    sockets = np.repeat(np.arange(n, dtype=np.int64), dv)
    for b in range(0, sockets.size, win):
        blk = sockets[b:b + win]
        sockets[b:b + win] = blk[rng.permutation(blk.size)]
    m = n * dv // dc
    gen = alist.Code(m, n, np.arange(0, m * dc + 1, dc, dtype=np.uint32), sockets.astype(np.uint32), 0)
    if code_arg[0] == "scrambled":  # the same code with its checks and variables renumbered at random: locality hidden
        vperm, cperm = rng.permutation(n), rng.permutation(m)
        cols = vperm[gen.col_idx.reshape(m, dc)[cperm]].reshape(-1)
        gen = alist.Code(m, n, gen.row_ptr, cols.astype(np.uint32), 0)
elif code_arg[0] == "seq":  # seq:log2n  (access-pattern experiment: a degenerate (3,6) multigraph whose variable-node
    n = 1 << int(code_arg[1])  # gathers are as sequential as the check-node rows: variable v owns edge rows 3v..3v+2)
    m = n // 2
    cols = np.repeat(np.arange(n, dtype=np.uint32), 3)
    gen = alist.Code(m, n, np.arange(0, 6 * m + 1, 6, dtype=np.uint32), cols, 0)
else:  # reg:log2n:dv:dc
    gen = alist.make_regular_code(1 << int(code_arg[1]), int(code_arg[2]), int(code_arg[3]), seed=1)
print("code", code_arg, "order", order_arg, "N", gen.n_vars, "M", gen.n_checks, "E", gen.n_edges, flush=True)
code = L.Code.from_generated(gen)
E, N, M = gen.n_edges, gen.n_vars, gen.n_checks
for cfg in configs:
    prec, log2p = cfg.split(":")
    log2p = int(log2p)
    P = 1 << log2p
    dt = np.float16 if prec == "fp16" else np.float32
    s = 2 if prec == "fp16" else 4
    dec = L.Decoder(code, L.FP16 if prec == "fp16" else L.FP32, log2p,
                    storage_order=L.ORDER_LOCALITY if order_arg == "locality" else L.ORDER_NATURAL)
    # P DISTINCT frames generated on the device (tiling a few frames over the slots would make lanes 16 apart look up
    # identical table entries and halve the shared-memory bank conflicts of the fp16 kernels: profiles/r06_ab_bulk_async.txt)
    import torch
    n_reg = N - gen.n_erased
    d_in = torch.empty((N, P), dtype=torch.float16 if prec == "fp16" else torch.float32, device="cuda")
    d_syn = torch.empty((P, (M + 31) // 32), dtype=torch.int32, device="cuda")
    d_ref = torch.empty((P, N // 32), dtype=torch.int32, device="cuda")
    dec.synthesize_frames(L.AWGN, 0.83, 0, P, d_in, d_syn, d_ref)
    dec.refill(L.AWGN, harness.channel_param("awgn", 0.83), d_in[:n_reg].cpu().numpy(), d_syn.cpu().numpy().view(np.uint32))
    del d_in, d_syn, d_ref
    dec.iterate(3)
    cn_ms, vn_ms = dec.iterate(20, timed=True)
    cn_b = (2 * E * s + 4 * ((M + 31) // 32)) * P
    vn_b = (2 * E + N) * s * P
    print(f"{cfg}: CN {cn_ms/20:.3f} ms {cn_b/(cn_ms/20*1e-3)/1e9:.0f} GB/s | VN {vn_ms/20:.3f} ms {vn_b/(vn_ms/20*1e-3)/1e9:.0f} GB/s | "
          f"pair {(cn_b+vn_b)/((cn_ms+vn_ms)/20*1e-3)/1e9:.0f} GB/s = {(cn_b+vn_b)/((cn_ms+vn_ms)/20*1e-3)/1e9/6551.4:.3f} of peak", flush=True)
    del dec
