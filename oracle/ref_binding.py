"""ctypes binding of oracle/_ref/libldpc_ref_{fp32,fp16}.so: the UNMODIFIED reference CUDA backend
compiled from /root/reference (oracle/Makefile `ref`, oracle/ref_driver.cu).  TEST INFRASTRUCTURE ONLY;
needs a GPU to run.  The two precisions are separate libraries (the reference picks llr_t at compile
time), so each is loaded in its own namespace."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(precision: str) -> str:
    return os.path.join(_HERE, "_ref", f"libldpc_ref_{precision}.so")


def available(precision: str = "fp32") -> bool:
    return os.path.exists(lib_path(precision))


_libs = {}


def _lib(precision: str):
    if precision not in _libs:
        L = C.CDLL(lib_path(precision), mode=os.RTLD_LOCAL)
        L.ref_last_error.restype = C.c_char_p
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_uint32]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_parallel_factor.restype = C.c_uint32
        L.ref_parallel_factor.argtypes = [C.c_void_p]
        L.ref_channel_param.restype = C.c_float
        L.ref_channel_param.argtypes = [C.c_void_p]
        L.ref_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.ref_refill.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.ref_backward.argtypes = [C.c_void_p]
        L.ref_forward.argtypes = [C.c_void_p, C.c_int]
        L.ref_check.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_deinterlace.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_read_messages.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_read_priors.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_graph.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _libs[precision] = L
    return _libs[precision]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefDecoder:
    """The reference's ldpc_decoder_gpu_cuda on an alist FILE (its ldpc_code only reads files)."""

    def __init__(self, alist_path: str, channel: str, noise: float, max_log2_frames: int, precision: str = "fp32"):
        self.L = _lib(precision)
        self.precision = precision
        self.np_dtype = np.float32 if precision == "fp32" else np.float16
        self._h = self.L.ref_create(alist_path.encode(), 1 if channel == "bsc" else 0, float(noise), max_log2_frames)
        if not self._h:
            raise RuntimeError(self.L.ref_last_error().decode())
        mne = np.zeros(6, dtype=np.int64)
        self.L.ref_graph(self._h, _p(mne), None, None, None, None)
        self.M, self.N, self.E, self.n_erased, self.max_deg_var, self.max_deg_check = (int(x) for x in mne)
        self.P = int(self.L.ref_parallel_factor(self._h))

    def channel_param(self) -> float:
        return float(self.L.ref_channel_param(self._h))

    def graph(self):
        row_ptr = np.zeros(self.M + 1, dtype=np.uint32)
        col_ptr = np.zeros(self.N + 1, dtype=np.uint32)
        in_to_out = np.zeros(self.E, dtype=np.uint32)
        edge_var = np.zeros(self.E, dtype=np.uint32)
        self.L.ref_graph(self._h, None, _p(row_ptr), _p(col_ptr), _p(in_to_out), _p(edge_var))
        return row_ptr, col_ptr, in_to_out, edge_var

    def decode(self, inputs: np.ndarray, syndromes: np.ndarray, num_iter_max=100, num_iter_check_parity=10):
        inputs = np.ascontiguousarray(inputs, dtype=self.np_dtype)
        syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
        n_frames = inputs.shape[1]
        results = np.zeros((n_frames, self.N // 32), dtype=np.uint32)
        out4 = np.zeros(4, dtype=np.float32)
        rc = self.L.ref_decode(self._h, num_iter_max, num_iter_check_parity, n_frames, _p(inputs), _p(syndromes),
                               _p(results), _p(out4))
        if rc:
            raise RuntimeError(self.L.ref_last_error().decode())
        return results, {"max_iter": int(out4[0]), "min_iter": int(out4[1]), "avg_iter": float(out4[2]),
                         "iter_time_per_vector": float(out4[3])}

    def refill(self, staged: np.ndarray, syndromes: np.ndarray):
        staged = np.ascontiguousarray(staged, dtype=self.np_dtype)
        syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
        assert self.L.ref_refill(self._h, staged.shape[1], _p(staged), _p(syndromes)) == 0

    def backward(self):
        assert self.L.ref_backward(self._h) == 0

    def forward(self, final_bits=False):
        assert self.L.ref_forward(self._h, int(final_bits)) == 0

    def check(self) -> np.ndarray:
        v = np.zeros(self.P, dtype=np.uint8)
        assert self.L.ref_check(self._h, _p(v)) == 0
        return v

    def deinterlace(self) -> np.ndarray:
        out = np.zeros((self.P, self.N // 32), dtype=np.uint32)
        assert self.L.ref_deinterlace(self._h, _p(out)) == 0
        return out

    def messages(self, slot: int) -> np.ndarray:
        out = np.zeros(self.E, dtype=np.float32)
        assert self.L.ref_read_messages(self._h, slot, _p(out)) == 0
        return out

    def priors(self, slot: int) -> np.ndarray:
        out = np.zeros(self.N, dtype=np.float32)
        assert self.L.ref_read_priors(self._h, slot, _p(out)) == 0
        return out

    def close(self):
        if getattr(self, "_h", None):
            self.L.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()
