"""alist tooling: code generators and a writer for the reference's alist dialect.

The reference ships no generator and its two sample codes are absent from the tree
(SURVEY.md section 8d "Missing code files"), so the harness builds stand-in codes here.

Dialect (reference src/ldpc_code.cpp:45-152): optional ``#e=<n_erased_vars>`` header
line; line 1 ``M N`` (checks first); line 2 ignored; line 3 the M check degrees; line 4
the N variable degrees; then M rows of 1-based variable indices.  The per-variable
adjacency section of a classic alist is never read, so it is not written.
"""
from __future__ import annotations

import numpy as np


class Code:
    """Check-major (CSR) description of a parity-check matrix."""

    def __init__(self, n_checks: int, n_vars: int, row_ptr: np.ndarray, col_idx: np.ndarray, n_erased: int = 0):
        self.n_checks = int(n_checks)
        self.n_vars = int(n_vars)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        self.col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
        self.n_erased = int(n_erased)
        assert self.row_ptr.shape == (self.n_checks + 1,)
        assert int(self.row_ptr[-1]) == self.col_idx.size

    @property
    def n_edges(self) -> int:
        return int(self.col_idx.size)

    def var_degrees(self) -> np.ndarray:
        return np.bincount(self.col_idx, minlength=self.n_vars).astype(np.int64)

    def check_degrees(self) -> np.ndarray:
        return np.diff(self.row_ptr.astype(np.int64))

    def csc(self):
        """(col_ptr[N+1], in_to_out[E]): in-edges of a variable ordered by increasing
        out-edge id, which is the order ldpc_code.cpp:139-151 assigns them."""
        order = np.argsort(self.col_idx, kind="stable").astype(np.uint32)
        col_ptr = np.zeros(self.n_vars + 1, dtype=np.uint32)
        np.cumsum(self.var_degrees(), out=col_ptr[1:])
        return col_ptr, order

    def syndrome_packed(self, bits_packed: np.ndarray) -> np.ndarray:
        """H*x for frame-major packed frames.  bits_packed: u32[F, N/32] LSB-first ->
        u32[F, ceil(M/32)] (same packing as reference main.cpp:535-537)."""
        bits_packed = np.ascontiguousarray(bits_packed, dtype=np.uint32)
        F = bits_packed.shape[0]
        bits = np.unpackbits(bits_packed.view(np.uint8), axis=1, bitorder="little")[:, : self.n_vars]
        # frame-sliced words: one column of packed frames per variable
        sliced = np.packbits(bits.T, axis=1, bitorder="little")  # [N, ceil(F/8)]
        gathered = sliced[self.col_idx]  # [E, F/8]
        deg = self.check_degrees()
        assert (deg > 0).all()
        syn = np.bitwise_xor.reduceat(gathered, self.row_ptr[:-1].astype(np.int64), axis=0)  # [M, F/8]
        syn_bits = np.unpackbits(syn, axis=1, bitorder="little")[:, :F].T  # [F, M]
        S = (self.n_checks + 31) // 32
        pad = np.zeros((F, S * 32), dtype=np.uint8)
        pad[:, : self.n_checks] = syn_bits
        return np.packbits(pad, axis=1, bitorder="little").view(np.uint32).reshape(F, S)


def write_alist(code: Code, path: str) -> None:
    deg_c = code.check_degrees()
    deg_v = code.var_degrees()
    with open(path, "w") as f:
        if code.n_erased:
            f.write(f"#e={code.n_erased}\n")
        f.write(f"{code.n_checks} {code.n_vars}\n")
        f.write(f"{int(deg_c.max())} {int(deg_v.max())}\n")
        f.write(" ".join(map(str, deg_c.tolist())) + "\n")
        f.write(" ".join(map(str, deg_v.tolist())) + "\n")
        cols = (code.col_idx.astype(np.int64) + 1).tolist()
        rp = code.row_ptr.tolist()
        lines = [" ".join(map(str, cols[rp[i] : rp[i + 1]])) for i in range(code.n_checks)]
        f.write("\n".join(lines))
        f.write("\n")


def alist_text(code: Code) -> str:
    import io
    import os
    import tempfile

    fd, p = tempfile.mkstemp(suffix=".alist")
    os.close(fd)
    try:
        write_alist(code, p)
        with io.open(p) as f:
            return f.read()
    finally:
        os.unlink(p)


def _match_sockets(var_of_socket: np.ndarray, check_deg: np.ndarray, rng: np.random.Generator, max_pass: int = 200):
    """Random socket matching without repeated (check, variable) pairs."""
    E = var_of_socket.size
    row_ptr = np.zeros(check_deg.size + 1, dtype=np.int64)
    np.cumsum(check_deg, out=row_ptr[1:])
    assert row_ptr[-1] == E
    check_of_slot = np.repeat(np.arange(check_deg.size, dtype=np.int64), check_deg)
    cols = var_of_socket[rng.permutation(E)].astype(np.int64)
    for _ in range(max_pass):
        key = check_of_slot * (cols.max() + 1) + cols
        order = np.argsort(key, kind="stable")
        dup = np.zeros(E, dtype=bool)
        dup[order[1:]] = key[order[1:]] == key[order[:-1]]
        idx = np.nonzero(dup)[0]
        if idx.size == 0:
            break
        # swap every offending slot with a random other slot
        other = rng.integers(0, E, size=idx.size)
        cols[idx], cols[other] = cols[other].copy(), cols[idx].copy()
    else:
        raise RuntimeError("could not remove repeated edges")
    # sort variables inside each check (cosmetic, deterministic)
    key = check_of_slot * (cols.max() + 1) + cols
    cols = cols[np.argsort(key, kind="stable")]
    return row_ptr.astype(np.uint32), cols.astype(np.uint32)


def _spread_check_degrees(n_edges: int, n_checks: int) -> np.ndarray:
    base = n_edges // n_checks
    deg = np.full(n_checks, base, dtype=np.int64)
    deg[: n_edges - base * n_checks] += 1
    return deg


def make_regular_code(n_vars: int, dv: int = 3, dc: int = 6, seed: int = 1) -> Code:
    """Random (dv,dc)-regular code, M = N*dv/dc checks (BASELINE.json config 5 shape)."""
    assert (n_vars * dv) % dc == 0 and n_vars % 32 == 0
    rng = np.random.default_rng(seed)
    n_checks = n_vars * dv // dc
    sockets = np.repeat(np.arange(n_vars, dtype=np.int64), dv)
    row_ptr, cols = _match_sockets(sockets, np.full(n_checks, dc, dtype=np.int64), rng)
    return Code(n_checks, n_vars, row_ptr, cols, 0)


def make_irregular_code(n_vars: int, n_checks: int, var_classes, n_erased: int = 0, seed: int = 1) -> Code:
    """Random irregular code.  var_classes: list of (count, degree) in variable-index
    order (the punctured variables must be the LAST n_erased indices: reference
    ldpc_decoder_gpu.cu:225).  Check degrees are concentrated on two adjacent values."""
    rng = np.random.default_rng(seed)
    counts = [int(c) for c, _ in var_classes]
    assert sum(counts) == n_vars and n_vars % 32 == 0
    deg = np.concatenate([np.full(c, d, dtype=np.int64) for c, d in var_classes])
    sockets = np.repeat(np.arange(n_vars, dtype=np.int64), deg)
    check_deg = _spread_check_degrees(int(deg.sum()), n_checks)
    row_ptr, cols = _match_sockets(sockets, check_deg, rng)
    return Code(n_checks, n_vars, row_ptr, cols, n_erased)


def make_readme_shaped_code(log2_n: int = 20, seed: int = 1) -> Code:
    """Stand-in for the missing code_awgn_rate_0.5_thr_0.95.alist (README.md:81-86):
    N = 2^log2_n variables, M = N*611669/2^20 checks, the last N*174763/2^20
    variables punctured, degrees <= 6, rate (N-M)/(N-e) = 0.5.

    Structure: the punctured variables have degree 6, 5/6 of the transmitted ones
    degree 3 and 1/6 degree 2 -- a random ensemble, NOT the optimised sample code, so
    its noise threshold is lower than 0.95."""
    n = 1 << log2_n
    m = int(round(n * 611669 / (1 << 20)))
    e = int(round(n * 174763 / (1 << 20)))
    n_tx = n - e
    n2 = n_tx // 6
    classes = [(n_tx - n2, 3), (n2, 2), (e, 6)]
    return make_irregular_code(n, m, classes, n_erased=e, seed=seed)


# ---------------------------------------------------------------------------------------------------------------
# Reordering for locality
# ---------------------------------------------------------------------------------------------------------------
def read_alist(path: str) -> Code:
    """Parser of the reference's dialect (src/ldpc_code.cpp:45-152), for the offline tools below: optional ``#e=`` /
    ``#ec=`` header lines, ``M N``, an ignored line, M check degrees, N variable degrees, M rows of 1-based variable
    indices padded with zeros.  (The decoder itself parses alists in C++: csrc/ldpc_code.cpp.)"""
    n_erased = 0
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while lines[i].startswith("#"):
        if lines[i].startswith("#e="):
            n_erased = int(lines[i][3:])
        i += 1
    m, n = (int(x) for x in lines[i].split()[:2])
    deg_c = np.array(lines[i + 2].split(), dtype=np.int64)
    assert deg_c.size == m and len(lines[i + 3].split()) == n
    cols = []
    for r in range(m):
        row = [int(x) for x in lines[i + 4 + r].split() if int(x) > 0]
        assert len(row) == deg_c[r], f"row {r}: {len(row)} entries, degree {deg_c[r]}"
        cols.extend(row)
    row_ptr = np.zeros(m + 1, dtype=np.uint32)
    np.cumsum(deg_c, out=row_ptr[1:])
    return Code(m, n, row_ptr, np.array(cols, dtype=np.int64) - 1, n_erased)


def gather_span(code: Code) -> np.ndarray:
    """Per variable: distance, in message rows (check-major edge ids), between the first and the last row its
    variable-node update gathers.  Small spans = the scattered 2-8 KB row reads of that pass stay close together
    (worth ~3 % of the pass on B200 at 2^12 frames in flight: profiles/r05_locality_experiment.txt)."""
    col_ptr, in_to_out = code.csc()
    first = in_to_out[col_ptr[:-1].astype(np.int64).clip(max=max(code.n_edges - 1, 0))].astype(np.int64)
    last = in_to_out[(col_ptr[1:].astype(np.int64) - 1).clip(min=0)].astype(np.int64)
    deg = np.diff(col_ptr.astype(np.int64))
    return np.where(deg > 0, last - first, 0)


def reorder_for_locality(code: Code):
    """Renumber checks and variables so that the edges of a variable sit in nearby message rows (reverse Cuthill-McKee
    on the bipartite graph).  The punctured variables stay the last ``n_erased`` indices, as the decoder requires
    (reference ldpc_decoder_gpu.cu:225).  Returns ``(new_code, var_perm, check_perm)`` with
    ``new variable j = old variable var_perm[j]`` and ``new check i = old check check_perm[i]``: feed the decoder
    ``input[var_perm]`` / syndrome bits ``[check_perm]`` and read decisions back through ``var_perm``
    (:func:`permute_frames`).  The result describes the same code up to this coordinate permutation; a random
    (expander-like) graph has no locality to find, structured ones (quasi-cyclic, spatially coupled, banded) do."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    m, n, e = code.n_checks, code.n_vars, code.n_edges
    rows = np.repeat(np.arange(m, dtype=np.int64), code.check_degrees())
    cols = code.col_idx.astype(np.int64) + m  # nodes: checks 0..m-1, variables m..m+n-1
    data = np.ones(2 * e, dtype=np.int8)
    adj = coo_matrix((data, (np.concatenate([rows, cols]), np.concatenate([cols, rows]))), shape=(m + n, m + n)).tocsr()
    order = np.asarray(reverse_cuthill_mckee(adj, symmetric_mode=True), dtype=np.int64)
    check_perm = order[order < m]
    var_order = order[order >= m] - m
    n_tx = n - code.n_erased
    var_perm = np.concatenate([var_order[var_order < n_tx], var_order[var_order >= n_tx]])  # punctured tail kept
    new_index_of_var = np.empty(n, dtype=np.int64)
    new_index_of_var[var_perm] = np.arange(n)
    deg = code.check_degrees()[check_perm]
    row_ptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    old_rp = code.row_ptr.astype(np.int64)
    # edges of new check i = edges of old check check_perm[i], variables renumbered and sorted inside the row
    src = np.concatenate([np.arange(old_rp[c], old_rp[c + 1]) for c in check_perm]) if m else np.zeros(0, np.int64)
    new_cols = new_index_of_var[code.col_idx.astype(np.int64)[src]]
    row_of = np.repeat(np.arange(m, dtype=np.int64), deg)
    new_cols = new_cols[np.lexsort((new_cols, row_of))]
    return Code(m, n, row_ptr.astype(np.uint32), new_cols.astype(np.uint32), code.n_erased), var_perm, check_perm


def permute_frames(code: Code, var_perm: np.ndarray, check_perm: np.ndarray, inp: np.ndarray, syn: np.ndarray):
    """Inputs of the ORIGINAL code -> inputs of the reordered one.  inp: [N or N - n_erased, F] channel values
    (element [i, v] = i-th value of frame v, the decoder's layout); syn: u32[F, ceil(M/32)] packed syndromes."""
    n_rows = inp.shape[0]
    new_inp = np.ascontiguousarray(inp[var_perm[:n_rows]])  # punctured tail stays last, so a truncated input still works
    bits = np.unpackbits(np.ascontiguousarray(syn, dtype=np.uint32).view(np.uint8), axis=1, bitorder="little")
    new_bits = np.zeros_like(bits)
    new_bits[:, : code.n_checks] = bits[:, check_perm]
    new_syn = np.packbits(new_bits, axis=1, bitorder="little").view(np.uint32).reshape(syn.shape)
    return new_inp, new_syn


def unpermute_results(var_perm: np.ndarray, results: np.ndarray) -> np.ndarray:
    """Packed decisions of the reordered code (u32[F, N/32], LSB first) -> variable order of the original code."""
    bits = np.unpackbits(np.ascontiguousarray(results, dtype=np.uint32).view(np.uint8), axis=1, bitorder="little")
    out = np.empty_like(bits)
    out[:, var_perm] = bits
    return np.packbits(out, axis=1, bitorder="little").view(np.uint32).reshape(results.shape)


if __name__ == "__main__":  # python -m ldpc_decoder_b200.alist reorder in.alist out.alist [perm.npz]
    import sys

    if len(sys.argv) >= 4 and sys.argv[1] == "reorder":
        c = read_alist(sys.argv[2])
        new, vp, cp = reorder_for_locality(c)
        write_alist(new, sys.argv[3])
        if len(sys.argv) > 4:
            np.savez(sys.argv[4], var_perm=vp, check_perm=cp)
        print(f"mean gather span {gather_span(c).mean():.0f} -> {gather_span(new).mean():.0f} rows "
              f"(N={c.n_vars}, M={c.n_checks}, E={c.n_edges})")
    else:
        print(__doc__)
