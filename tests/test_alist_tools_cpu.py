"""CPU tests of the offline alist tools: reader/writer round trip and reordering for locality."""
import numpy as np

from ldpc_decoder_b200 import alist, harness


def _banded_code(n=4096, dv=3, dc=6, win=96, seed=2):
    """(dv,dc) code whose variable sockets are shuffled only inside windows of `win` sockets: every variable's
    checks are close to each other (the structure reordering should recover)."""
    rng = np.random.default_rng(seed)
    sockets = np.repeat(np.arange(n, dtype=np.int64), dv)
    for b in range(0, sockets.size, win):
        blk = sockets[b:b + win]
        sockets[b:b + win] = blk[rng.permutation(blk.size)]
    m = n * dv // dc
    rows = [np.unique(sockets[c * dc:(c + 1) * dc]) for c in range(m)]  # drop repeated (check, variable) pairs
    row_ptr = np.zeros(m + 1, dtype=np.uint32)
    np.cumsum([r.size for r in rows], out=row_ptr[1:])
    return alist.Code(m, n, row_ptr, np.concatenate(rows).astype(np.uint32), 0)


def _scramble(code, seed=3, keep_tail=0):
    """Random renumbering of checks and variables (the last keep_tail variables stay last)."""
    rng = np.random.default_rng(seed)
    n_tx = code.n_vars - keep_tail
    vp = np.concatenate([rng.permutation(n_tx), np.arange(n_tx, code.n_vars)])
    cp = rng.permutation(code.n_checks)
    new_of_var = np.empty(code.n_vars, dtype=np.int64)
    new_of_var[vp] = np.arange(code.n_vars)
    rp = code.row_ptr.astype(np.int64)
    rows = [np.sort(new_of_var[code.col_idx[rp[c]:rp[c + 1]].astype(np.int64)]) for c in cp]
    row_ptr = np.zeros(code.n_checks + 1, dtype=np.uint32)
    np.cumsum([r.size for r in rows], out=row_ptr[1:])
    return alist.Code(code.n_checks, code.n_vars, row_ptr, np.concatenate(rows).astype(np.uint32), keep_tail)


def test_read_alist_round_trip(tmp_path):
    code = alist.make_readme_shaped_code(10, seed=3)
    p = tmp_path / "c.alist"
    alist.write_alist(code, str(p))
    back = alist.read_alist(str(p))
    assert (back.n_checks, back.n_vars, back.n_erased) == (code.n_checks, code.n_vars, code.n_erased)
    assert np.array_equal(back.row_ptr, code.row_ptr) and np.array_equal(back.col_idx, code.col_idx)


def test_reordering_is_a_coordinate_permutation_of_the_same_code():
    code = alist.make_readme_shaped_code(11, seed=4)  # punctured tail
    new, vp, cp = alist.reorder_for_locality(code)
    assert sorted(vp.tolist()) == list(range(code.n_vars)) and sorted(cp.tolist()) == list(range(code.n_checks))
    n_tx = code.n_vars - code.n_erased
    assert new.n_erased == code.n_erased and set(vp[n_tx:].tolist()) == set(range(n_tx, code.n_vars))
    assert np.array_equal(np.sort(new.check_degrees()), np.sort(code.check_degrees()))
    # H'[i, j] = H[check_perm[i], var_perm[j]]: syndromes of permuted words are the permuted syndromes
    inp, syn, bits = harness.make_frames(code, 40, "awgn", 0.7, seed=5, dtype=np.float32)
    new_inp, new_syn = alist.permute_frames(code, vp, cp, inp, syn)
    bits = np.ascontiguousarray(bits)
    b = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")
    new_bits = np.ascontiguousarray(np.packbits(np.ascontiguousarray(b[:, vp]), axis=1, bitorder="little")).view(np.uint32).reshape(bits.shape)
    assert np.array_equal(new.syndrome_packed(new_bits), new_syn)
    assert np.array_equal(new_inp, inp[vp[:inp.shape[0]]])
    assert np.array_equal(alist.unpermute_results(vp, new_bits), bits)


def test_reordering_recovers_hidden_locality():
    banded = _banded_code()
    scrambled = _scramble(banded)
    new, _, _ = alist.reorder_for_locality(scrambled)
    span_banded = alist.gather_span(banded).mean()
    span_scrambled = alist.gather_span(scrambled).mean()
    span_new = alist.gather_span(new).mean()
    assert span_scrambled > 20 * span_banded          # the scramble destroyed the structure ...
    assert span_new < span_scrambled / 10             # ... and the reordering finds most of it again
    # a random graph has none to find: the tool must still return a valid code, not a better one
    rnd = alist.make_regular_code(2048, 3, 6, seed=9)
    new_rnd, vp, cp = alist.reorder_for_locality(rnd)
    assert new_rnd.n_edges == rnd.n_edges and alist.gather_span(new_rnd).mean() > 0.2 * alist.gather_span(rnd).mean()
