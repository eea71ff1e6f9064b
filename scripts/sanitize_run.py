"""Small decodes that walk every kernel path, to be run under compute-sanitizer on the GPU box:
  compute-sanitizer --tool memcheck  python scripts/sanitize_run.py
  compute-sanitizer --tool racecheck python scripts/sanitize_run.py bulk     (shared-memory hazards of the bulk kernels)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness

only_bulk = len(sys.argv) > 1 and sys.argv[1] == "bulk"
gen = alist.make_readme_shaped_code(11, seed=5)
code = L.Code.from_generated(gen)
hide = alist.make_irregular_code(512, 64, [(448, 3), (32, 10), (32, 12)], n_erased=0, seed=7)
for prec, dt in ((L.FP16, np.float16), (L.FP32, np.float32)):
    inp, syn, bits = harness.make_frames(gen, 150, "awgn", 0.6, seed=3, dtype=dt)
    prm = harness.channel_param("awgn", 0.6)
    if only_bulk:
        os.environ["LDPC_B200_CN_BULK"] = "1"
        os.environ["LDPC_B200_VN_BULK"] = "1"
        log2p = 11 if prec == L.FP16 else 10
        dec = L.Decoder(code, prec, log2p)
        n = (1 << log2p) + 300
        r, it, rep = dec.decode(L.AWGN, prm, inp, syn, n_frames=n, num_iter_max=30, pool_frames=150, flags=L.FLAG_EARLY_RETIRE)
        print("bulk", prec, rep["avg_iter"], flush=True)
        continue
    dec = L.Decoder(code, prec, 5)
    r0, it0, _ = dec.decode(L.AWGN, prm, inp, syn, num_iter_max=40)
    for flags in (L.FLAG_NO_GRAPH, L.FLAG_EARLY_RETIRE, L.FLAG_TIME_KERNELS):
        r, it, _ = dec.decode(L.AWGN, prm, inp, syn, num_iter_max=40, flags=flags)
        assert np.array_equal(it, it0)
    loc = L.Decoder(code, prec, 5, storage_order=L.ORDER_LOCALITY)
    r, it, _ = loc.decode(L.AWGN, prm, inp, syn, num_iter_max=40)
    assert np.array_equal(r, r0) and np.array_equal(it, it0)
    q = L.FrameQueue(150, 7)
    res = np.zeros_like(r0)
    it, rep = dec.decode_from_queue(q, L.AWGN, prm, inp, syn, res, num_iter_max=40)
    assert np.array_equal(res, r0)
    rm, itm, repm, per = L.decode_multi([dec, loc], L.AWGN, prm, inp, syn, num_iter_max=40, block_frames=16)
    assert repm["frames_decoded"] == 150
    if prec == L.FP32:
        r, it, _ = dec.decode(L.AWGN, prm, inp.astype(np.float16), syn, num_iter_max=40, flags=L.FLAG_HALF_INPUT)
    os.environ["LDPC_B200_FUSED_REFILL"] = "0"
    plain = L.Decoder(code, prec, 5)
    del os.environ["LDPC_B200_FUSED_REFILL"]
    r, it, _ = plain.decode(L.AWGN, prm, inp, syn, num_iter_max=40)
    assert np.array_equal(r, r0)
    hi = L.Decoder(L.Code.from_generated(hide), prec, 5)
    hinp, hsyn, _ = harness.make_frames(hide, 80, "bsc", 0.01, seed=4, dtype=dt)
    hi.decode(L.BSC, harness.channel_param("bsc", 0.01), hinp, hsyn, num_iter_max=20, num_iter_check_parity=4)
    os.environ["LDPC_B200_IDX64"] = "1"
    d64 = L.Decoder(code, prec, 5)
    del os.environ["LDPC_B200_IDX64"]
    r, it, _ = d64.decode(L.AWGN, prm, inp, syn, num_iter_max=40)
    assert np.array_equal(r, r0)
    print("ok", prec, flush=True)
print("done")
