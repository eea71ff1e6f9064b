"""ctypes binding of oracle/_ref/libldpc_oracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libldpc_oracle.so")

AWGN, BSC = 0, 1
FP32, FP16 = 0, 1
PRE_THRESHOLD_CUDA = np.float32(1e-5)  # flood.cu:14
PRE_THRESHOLD_OPENCL = np.float32(2.0 * np.exp(-11.0))  # flood_vec2.cl:187 with threshold 10


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("ldpc_oracle.c", "ldpc_oracle.h", "ldpc_cpu_fast.c", "ldpc_oracle_internal.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


class _Code(C.Structure):
    _fields_ = [
        ("n_checks", C.c_int64),
        ("n_vars", C.c_int64),
        ("n_edges", C.c_int64),
        ("n_erased_vars", C.c_int64),
        ("n_erased_checks", C.c_int64),
        ("max_deg_var", C.c_int32),
        ("max_deg_check", C.c_int32),
        ("out_bit_to_edge", C.POINTER(C.c_uint32)),
        ("in_bit_to_edge", C.POINTER(C.c_uint32)),
        ("edge_out_to_in", C.POINTER(C.c_uint32)),
        ("edge_in_to_out", C.POINTER(C.c_uint32)),
        ("in_edge_to_bit", C.POINTER(C.c_uint32)),
        ("out_edge_to_bit", C.POINTER(C.c_uint32)),
        ("out_edge_to_in_bit", C.POINTER(C.c_uint32)),
    ]


class _Report(C.Structure):
    _fields_ = [
        ("max_iter", C.c_uint32),
        ("min_iter", C.c_uint32),
        ("avg_iter", C.c_float),
        ("iter_time_per_vector", C.c_float),
        ("global_iters", C.c_uint32),
        ("elapsed_s", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_code_from_alist_file.restype = C.POINTER(_Code)
        L.oracle_code_from_alist_file.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.oracle_code_from_alist_text.restype = C.POINTER(_Code)
        L.oracle_code_from_alist_text.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.oracle_code_free.argtypes = [C.POINTER(_Code)]
        L.oracle_code_rate.restype = C.c_float
        L.oracle_code_rate.argtypes = [C.POINTER(_Code)]
        L.oracle_compute_syndrome.argtypes = [C.POINTER(_Code), C.c_void_p, C.c_void_p]
        L.oracle_phi_abs.restype = C.c_float
        L.oracle_phi_abs.argtypes = [C.c_float, C.c_float]
        L.oracle_phi.restype = C.c_float
        L.oracle_phi.argtypes = [C.c_float, C.c_float]
        L.oracle_phi_abs_h.restype = C.c_uint16
        L.oracle_phi_abs_h.argtypes = [C.c_uint16]
        L.oracle_phi_table_h.argtypes = [C.c_void_p]
        L.oracle_decoder_create.restype = C.c_void_p
        L.oracle_decoder_create.argtypes = [C.POINTER(_Code), C.c_int, C.c_uint32, C.c_float, C.c_void_p, C.c_int]
        L.oracle_decoder_free.argtypes = [C.c_void_p]
        L.oracle_decoder_set_fast.argtypes = [C.c_void_p, C.c_int]
        L.oracle_refill.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_backward.argtypes = [C.c_void_p]
        L.oracle_forward.argtypes = [C.c_void_p, C.c_int]
        L.oracle_check_parity.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_permute.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_deinterlace.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_read_messages.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_read_priors.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_decode.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(_Report), C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleCode:
    def __init__(self, path: str | None = None, text: str | None = None):
        err = C.c_int(0)
        if path is not None:
            self._c = lib().oracle_code_from_alist_file(path.encode(), C.byref(err))
        else:
            self._c = lib().oracle_code_from_alist_text(text.encode(), C.byref(err))
        if not self._c:
            raise ValueError({1: "Alist file could not be opened for reading", 2: "malformed alist file"}.get(err.value, "error"))
        c = self._c.contents
        self.n_checks, self.n_vars, self.n_edges = c.n_checks, c.n_vars, c.n_edges
        self.n_erased = c.n_erased_vars
        self.n_erased_checks = c.n_erased_checks
        self.max_deg_var, self.max_deg_check = c.max_deg_var, c.max_deg_check

    def _arr(self, name, n):
        return np.ctypeslib.as_array(getattr(self._c.contents, name), shape=(n,)).copy()

    @property
    def row_ptr(self):
        return self._arr("out_bit_to_edge", self.n_checks + 1)

    @property
    def col_ptr(self):
        return self._arr("in_bit_to_edge", self.n_vars + 1)

    @property
    def in_to_out(self):
        return self._arr("edge_in_to_out", self.n_edges)

    @property
    def out_to_in(self):
        return self._arr("edge_out_to_in", self.n_edges)

    @property
    def out_edge_to_var(self):
        return self._arr("out_edge_to_in_bit", self.n_edges)

    def rate(self):
        return lib().oracle_code_rate(self._c)

    def syndrome(self, bits_packed: np.ndarray) -> np.ndarray:
        bits_packed = np.ascontiguousarray(bits_packed, dtype=np.uint32)
        S = (self.n_checks + 31) // 32
        out = np.zeros((bits_packed.shape[0], S), dtype=np.uint32)
        for f in range(bits_packed.shape[0]):
            lib().oracle_compute_syndrome(self._c, _p(bits_packed[f]), _p(out[f]))
        return out

    def __del__(self):
        if getattr(self, "_c", None):
            lib().oracle_code_free(self._c)
            self._c = None


def phi_abs(x, pre=PRE_THRESHOLD_CUDA):
    return np.array([lib().oracle_phi_abs(float(v), float(pre)) for v in np.atleast_1d(x)], dtype=np.float32)


def phi_table_h() -> np.ndarray:
    t = np.zeros(32768, dtype=np.uint16)
    lib().oracle_phi_table_h(_p(t))
    return t


class OracleDecoder:
    def __init__(self, code: OracleCode, precision=FP32, log2_frames=5, pre_threshold=PRE_THRESHOLD_CUDA,
                 phi_table=None, n_threads=None, fast=False):
        self.code = code
        self.precision = precision
        self.P = 1 << log2_frames
        self.np_dtype = np.float32 if precision == FP32 else np.float16
        if n_threads is None:
            n_threads = os.cpu_count() or 1
        tab = None
        if phi_table is not None:
            tab = np.ascontiguousarray(phi_table, dtype=np.uint16)
            assert tab.size == 32768
        self._d = lib().oracle_decoder_create(code._c, precision, log2_frames, float(pre_threshold),
                                              _p(tab) if tab is not None else None, n_threads)
        if not self._d:
            raise ValueError("This decoder only handles input sizes that are multiple of 32")
        if fast:  # SIMD-over-frames CN/VN kernels (CPU baseline timing; fp32 only)
            lib().oracle_decoder_set_fast(self._d, 1)

    def refill(self, channel, param, staged: np.ndarray, syndromes: np.ndarray):
        """staged: [n_regular, cnt] channel values (frame fastest); syndromes [cnt, S]."""
        staged = np.ascontiguousarray(staged, dtype=self.np_dtype)
        syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
        cnt = staged.shape[1]
        lib().oracle_refill(self._d, channel, float(param), cnt, _p(staged), _p(syndromes))

    def backward(self):
        lib().oracle_backward(self._d)

    def forward(self, final_bits=False):
        lib().oracle_forward(self._d, int(final_bits))

    def check_parity(self) -> np.ndarray:
        v = np.zeros(self.P, dtype=np.uint8)
        lib().oracle_check_parity(self._d, _p(v))
        return v

    def deinterlace(self) -> np.ndarray:
        out = np.zeros((self.P, self.code.n_vars // 32), dtype=np.uint32)
        lib().oracle_deinterlace(self._d, _p(out))
        return out

    def messages(self, slot: int) -> np.ndarray:
        out = np.zeros(self.code.n_edges, dtype=np.float32)
        lib().oracle_read_messages(self._d, slot, _p(out))
        return out

    def priors(self, slot: int) -> np.ndarray:
        out = np.zeros(self.code.n_vars, dtype=np.float32)
        lib().oracle_read_priors(self._d, slot, _p(out))
        return out

    def decode(self, channel, param, num_iter_max, num_iter_check_parity, inputs: np.ndarray, syndromes: np.ndarray,
               simd_pairs=False):
        """inputs: [N (or n_regular), n_frames] reference layout; returns (results, iter_counts, report dict)."""
        inputs = np.ascontiguousarray(inputs, dtype=self.np_dtype)
        syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
        n_frames = inputs.shape[1]
        results = np.zeros((n_frames, self.code.n_vars // 32), dtype=np.uint32)
        iters = np.zeros(n_frames, dtype=np.uint32)
        rep = _Report()
        lib().oracle_decode(self._d, channel, float(param), num_iter_max, num_iter_check_parity, n_frames, _p(inputs),
                            _p(syndromes), _p(results), C.byref(rep), _p(iters), int(simd_pairs))
        return results, iters, {k: getattr(rep, k) for k, _ in _Report._fields_}

    def __del__(self):
        if getattr(self, "_d", None):
            lib().oracle_decoder_free(self._d)
            self._d = None
