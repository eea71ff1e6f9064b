"""Diagnostic: per-iteration fp32 message error of (ours, reference build, CPU oracle) pairwise."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness
from oracle import binding as ob, ref_binding as rb

def stats(a, b):
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
    return f"max {rel.max():.2e} p99.9 {np.percentile(rel,99.9):.2e} frac>1e-3 {(rel>1e-3).mean():.2e} l2 {np.linalg.norm(a-b)/np.linalg.norm(b):.2e}"

for name, gen, sigma in (("reg36", alist.make_regular_code(2048,3,6,seed=3), 0.8), ("readme13", alist.make_readme_shaped_code(13, seed=5), 0.8), ("readme13@.55", alist.make_readme_shaped_code(13, seed=5), 0.55)):
    d = tempfile.mkdtemp(); path = os.path.join(d, "c.alist"); alist.write_alist(gen, path)
    P = 32
    inp, syn, bits = harness.make_frames(gen, P, "awgn", sigma, seed=1)
    prm = harness.channel_param("awgn", sigma)
    n_reg = gen.n_vars - gen.n_erased
    dec = L.Decoder(L.Code(path=path), L.FP32, 5)
    od = ob.OracleDecoder(ob.OracleCode(path=path), ob.FP32, 5)
    ref = rb.RefDecoder(path, "awgn", sigma, 5, "fp32")
    dec.refill(L.AWGN, prm, inp[:n_reg], syn); od.refill(ob.AWGN, prm, inp[:n_reg], syn); ref.refill(inp[:n_reg], syn)
    print("==", name)
    for it in range(1, 21):
        dec.iterate(1, True); od.backward(); od.forward(True); ref.backward(); ref.forward(True)
        if it in (1,2,3,5,10,20):
            a = np.concatenate([dec.messages(s) for s in (0,7,31)]); b = np.concatenate([od.messages(s) for s in (0,7,31)]); c = np.concatenate([ref.messages(s) for s in (0,7,31)])
            print(it, "ours-vs-ref   ", stats(a, c))
            print(it, "ours-vs-oracle", stats(a, b))
            print(it, "ref-vs-oracle ", stats(c, b))
            h = ref.deinterlace(); ho = od.deinterlace()
            print(it, "hard diffs ours/ref, oracle/ref:", sum(int((dec.hard(s)!=h[s]).sum()) for s in range(32)), int((ho!=h).sum()), "viol", int(ref.check().sum()))
