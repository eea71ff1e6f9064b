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
