"""FER / BER / iteration-histogram agreement with the reference's own sm_100 build (north_star parity item 3).
Same code, same frames, near-threshold noise so that a sizeable fraction of frames fails or runs long."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness
from oracle import ref_binding as rb


def wilson(k, n, z=1.96):
    if n == 0:
        return (0.0, 1.0)
    p = k / n
    d = 1 + z * z / n
    c = p + z * z / (2 * n)
    h = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n))
    return (float((c - h) / d), float((c + h) / d))


def run(log2n=15, n_frames=1024, sigma=0.805, imax=120, log2p=8, out=None):
    gen = alist.make_readme_shaped_code(log2n, seed=3)
    d = tempfile.mkdtemp(); path = os.path.join(d, "c.alist"); alist.write_alist(gen, path)
    res_all = {}
    for prec in ("fp32", "fp16"):
        dt = np.float16 if prec == "fp16" else np.float32
        inp, syn, bits = harness.make_frames(gen, n_frames, "awgn", sigma, seed=17, dtype=dt)
        prm = harness.channel_param("awgn", sigma)
        ref = rb.RefDecoder(path, "awgn", sigma, log2p, prec)
        rres, rrep = ref.decode(inp, syn, imax, 10)
        ref.close()
        dec = L.Decoder(L.Code(path=path), L.FP16 if prec == "fp16" else L.FP32, log2p)
        res, iters, rep = dec.decode(L.AWGN, prm, inp, syn, num_iter_max=imax)
        e_new = harness.count_bit_errors(res, bits); e_ref = harness.count_bit_errors(rres, bits)
        fe_new, fe_ref = int((e_new > 0).sum()), int((e_ref > 0).sum())
        row = {
            "n_frames": n_frames, "sigma": sigma, "code": {"N": gen.n_vars, "M": gen.n_checks, "E": gen.n_edges},
            "fer_new": fe_new / n_frames, "fer_ref": fe_ref / n_frames,
            "fer_ci95_new": wilson(fe_new, n_frames), "fer_ci95_ref": wilson(fe_ref, n_frames),
            "ber_new": float(e_new.sum() / (n_frames * gen.n_vars)), "ber_ref": float(e_ref.sum() / (n_frames * gen.n_vars)),
            "frames_with_identical_decisions": float((res == rres).all(axis=1).mean()),
            "frames_failing_in_exactly_one_decoder": int(((e_new > 0) != (e_ref > 0)).sum()),
            "iter_new": {"avg": rep["avg_iter"], "min": rep["min_iter"], "max": rep["max_iter"]},
            "iter_ref": {"avg": rrep["avg_iter"], "min": rrep["min_iter"], "max": rrep["max_iter"]},
            "iter_hist_new": {int(k): int(v) for k, v in zip(*np.unique(iters, return_counts=True))},
        }
        res_all[prec] = row
        print(prec, json.dumps({k: v for k, v in row.items() if k != "iter_hist_new"}), flush=True)
    if out:
        json.dump(res_all, open(out, "w"), indent=1)
    return res_all


if __name__ == "__main__":
    allr = {}
    for sigma in (0.805, 0.82, 0.83, 0.84):
        allr[str(sigma)] = run(n_frames=2048, sigma=sigma)
    json.dump(allr, open("gpurun_out/r06_fer_compare.json", "w"), indent=1)
