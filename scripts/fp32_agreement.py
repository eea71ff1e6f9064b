"""How closely do fp32 decodes agree?  Per-frame iteration counts and decisions of this decoder against (a) the CPU
oracle (libm phi) and (b) the reference's own sm_100 build, on the decode() cases of tests/test_gpu_parity.py plus larger
samples.  The numbers set the thresholds of the fp32 assertions."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness
from oracle import binding as ob
from oracle import ref_binding as rb

codes = {
    "reg36": alist.make_regular_code(2048, 3, 6, seed=3),
    "readme13": alist.make_readme_shaped_code(13, seed=5),
    "hideg": alist.make_irregular_code(1024, 128, [(896, 3), (64, 10), (64, 12)], n_erased=0, seed=7),
    "readme15": alist.make_readme_shaped_code(15, seed=1),
}
cases = [("reg36", "awgn", 0.78, 2000, 5, 60, 10), ("reg36", "awgn", 0.95, 400, 3, 20, 5), ("readme13", "awgn", 0.55, 1000, 4, 100, 10),
         ("hideg", "bsc", 0.01, 1000, 5, 40, 4), ("readme13", "awgn", 0.6, 1000, 6, 60, 10), ("readme15", "awgn", 0.80, 1024, 8, 120, 10),
         ("readme15", "awgn", 0.83, 1024, 8, 120, 10)]
out = []
d = tempfile.mkdtemp()
for name, channel, noise, n, log2p, imax, k in cases:
    gen = codes[name]
    path = os.path.join(d, name + ".alist")
    alist.write_alist(gen, path)
    inp, syn, bits = harness.make_frames(gen, n, channel, noise, seed=5, dtype=np.float32)
    prm = harness.channel_param(channel, noise)
    ch = L.AWGN if channel == "awgn" else L.BSC
    res, iters, rep = L.Decoder(L.Code(path=path), L.FP32, log2p).decode(ch, prm, inp, syn, num_iter_max=imax, num_iter_check_parity=k)
    row = {"case": [name, channel, noise, n, log2p, imax, k], "frac_converged": float((iters < imax).mean())}
    if gen.n_vars <= 8192:
        ores, oiters, _ = ob.OracleDecoder(ob.OracleCode(path=path), ob.FP32, log2p).decode(ob.AWGN if channel == "awgn" else ob.BSC, prm, imax, k, inp, syn)
        same = iters == oiters
        conv = same & (iters < imax)
        row["vs_oracle"] = {"iter_counts_equal": float(same.mean()), "decisions_equal_on_converged": bool(np.array_equal(res[conv], ores[conv])),
                            "frames_with_equal_decisions": float((res == ores).all(axis=1).mean())}
    if rb.available("fp32") and not (channel == "bsc" and gen.n_erased):
        ref = rb.RefDecoder(path, channel, noise, log2p, "fp32")
        rres, rrep = ref.decode(inp, syn, imax, k)
        ref.close()
        conv = iters < imax
        row["vs_reference_build"] = {"frames_with_equal_decisions": float((res == rres).all(axis=1).mean()),
                                     "equal_decisions_among_converged": float((res[conv] == rres[conv]).all(axis=1).mean()) if conv.any() else None,
                                     "avg_iter": [rep["avg_iter"], rrep["avg_iter"]], "max_min": [[rep["max_iter"], rep["min_iter"]], [rrep["max_iter"], rrep["min_iter"]]]}
    out.append(row)
    print(json.dumps(row), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r06_fp32_agreement.json"), "w"), indent=1)
