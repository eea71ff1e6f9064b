"""B200-native LDPC flood decoder: Python host binding over the C ABI (include/ldpc_b200.h).

The product is the CUDA library; this module only marshals numpy / torch buffers into it.
Names follow the reference's decoder API (h/ldpc_decoder_gpu.h:77-128):
``Decoder(code, precision, max_log2_frames)`` ~ the constructor, ``decode`` ~ decode(),
``parallel_factor`` ~ parallel_factor().
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import AWGN, BSC, LLR, FP16, FP32, LAYOUT_FRAME_MAJOR, LAYOUT_INTERLEAVED, LAYOUT_PACKED_BITS, FLAG_SIMD_PAIRS, FLAG_TIME_KERNELS, FLAG_NO_GRAPH, FLAG_EARLY_RETIRE, FLAG_HALF_INPUT, ORDER_NATURAL, ORDER_LOCALITY  # noqa: F401


class LdpcError(RuntimeError):
    """Mirror of the reference's `error` exception (h/common.h:61-73)."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def _check(status: int):
    if status != _lib.OK:
        raise LdpcError(status, _lib.lib().ldpc_b200_last_error().decode(errors="replace"))


def _ptr(a):
    """numpy array, torch tensor (host or CUDA) or raw int address -> void*."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _wait_for_torch(*buffers):
    """The library works on the decoder's own (non-blocking) streams; a CUDA tensor torch is still filling on ITS
    stream (torch.zeros, a copy) must be complete before the library touches it -- the C ABI leaves that to the caller."""
    for b in buffers:
        if getattr(b, "is_cuda", False):
            import torch

            torch.cuda.current_stream(b.device).synchronize()


class Code:
    """Code graph (reference ldpc_code, h/ldpc_code.h:10-54)."""

    def __init__(self, path: str | None = None, text: str | None = None, csr=None, n_erased: int = 0):
        L = _lib.lib()
        h = C.c_void_p()
        if path is not None:
            _check(L.ldpc_b200_code_load_alist(path.encode(), C.byref(h)))
        elif text is not None:
            b = text.encode()
            _check(L.ldpc_b200_code_parse_alist(b, len(b), C.byref(h)))
        else:
            n_checks, n_vars, row_ptr, col_idx = csr
            row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
            col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
            _check(L.ldpc_b200_code_from_csr(n_checks, n_vars, _ptr(row_ptr), _ptr(col_idx), n_erased, C.byref(h)))
        self._h = h
        info = _lib.CodeInfo()
        _check(L.ldpc_b200_code_get_info(self._h, C.byref(info)))
        self._info = info
        self.n_checks, self.n_vars, self.n_edges = info.n_checks, info.n_vars, info.n_edges
        self.n_erased = info.n_erased_vars
        self.n_erased_checks = info.n_erased_checks
        self.max_deg_var, self.max_deg_check = info.max_deg_var, info.max_deg_check

    @classmethod
    def from_generated(cls, gen):
        """gen: ldpc_decoder_b200.alist.Code"""
        return cls(csr=(gen.n_checks, gen.n_vars, gen.row_ptr, gen.col_idx), n_erased=gen.n_erased)

    def _arr(self, name, n):
        return np.ctypeslib.as_array(getattr(self._info, name), shape=(n,)).copy()

    @property
    def row_ptr(self):
        return self._arr("row_ptr", self.n_checks + 1)

    @property
    def col_ptr(self):
        return self._arr("col_ptr", self.n_vars + 1)

    @property
    def in_to_out(self):
        return self._arr("in_to_out", self.n_edges)

    @property
    def out_to_in(self):
        return self._arr("out_to_in", self.n_edges)

    @property
    def out_edge_to_var(self):
        return self._arr("out_edge_to_var", self.n_edges)

    def locality_order(self):
        """(check_order[M], var_order[N]): the storage order a Decoder built with storage_order=ORDER_LOCALITY uses."""
        co = np.zeros(self.n_checks, dtype=np.uint32)
        vo = np.zeros(self.n_vars, dtype=np.uint32)
        _check(_lib.lib().ldpc_b200_code_locality_order(self._h, _ptr(co), _ptr(vo)))
        return co, vo

    def rate(self) -> float:
        return float(np.float32(self.n_vars - self.n_checks) / np.float32(self.n_vars - self.n_erased))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:  # module globals may be gone at interpreter shutdown
            _lib.lib().ldpc_b200_code_destroy(self._h)
            self._h = None


class Decoder:
    """The decoder object (reference ldpc_decoder_gpu / ldpc_decoder_gpu_cuda)."""

    def __init__(self, code: Code, precision: int = FP32, max_log2_frames: int = 5, device: int = -1,
                 mem_headroom_pct: int = 0, storage_order: int = ORDER_NATURAL):
        cfg = _lib.Config(precision=precision, max_log2_frames=max_log2_frames, device=device,
                          mem_headroom_pct=mem_headroom_pct, storage_order=storage_order)
        h = C.c_void_p()
        _check(_lib.lib().ldpc_b200_decoder_create(code._h, C.byref(cfg), C.byref(h)))
        self._h = h
        self.code = code
        self.precision = precision
        self.np_dtype = np.float32 if precision == FP32 else np.float16

    def parallel_factor(self) -> int:
        return int(_lib.lib().ldpc_b200_decoder_frames_in_flight(self._h))

    def prepare_host_io(self):
        _check(_lib.lib().ldpc_b200_decoder_prepare_host_io(self._h))

    def set_erased_variables(self, n: int):
        _lib.lib().ldpc_b200_decoder_set_erased_vars(self._h, n)

    def bytes_per_frame_iteration(self) -> int:
        return int(_lib.lib().ldpc_b200_bytes_per_frame_iteration(self._h))

    def decode(self, channel: int, channel_param: float, inputs, syndromes, n_frames: int | None = None,
               num_iter_max: int = 100, num_iter_check_parity: int = 10, results=None, layout=LAYOUT_INTERLEAVED,
               pool_frames: int = 0, flags: int = 0, frame_stride: int = 0):
        """inputs/syndromes/results: numpy arrays (host) or torch CUDA tensors (device resident).
        Returns (results, iter_counts, report dict)."""
        if isinstance(inputs, np.ndarray):
            in_dtype = np.uint32 if layout == LAYOUT_PACKED_BITS else (np.float16 if flags & FLAG_HALF_INPUT else self.np_dtype)
            inputs = np.ascontiguousarray(inputs, dtype=in_dtype)
            if n_frames is None:
                n_frames = inputs.shape[1] if layout == LAYOUT_INTERLEAVED else inputs.shape[0]
        if isinstance(syndromes, np.ndarray):
            syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
        assert n_frames is not None
        if results is None:
            results = np.zeros((n_frames, self.code.n_vars // 32), dtype=np.uint32)
        iters = np.zeros(n_frames, dtype=np.uint32)
        prm = _lib.DecodeParams(channel=channel, channel_param=channel_param, num_iter_max=num_iter_max,
                                num_iter_check_parity=num_iter_check_parity, input_layout=layout,
                                frame_stride=frame_stride, input_pool_frames=pool_frames, flags=flags)
        rep = _lib.Report()
        _wait_for_torch(inputs, syndromes, results)
        _check(_lib.lib().ldpc_b200_decode(self._h, C.byref(prm), n_frames, _ptr(inputs), _ptr(syndromes),
                                           _ptr(results), C.byref(rep), _ptr(iters)))
        return results, iters, {k: getattr(rep, k) for k, _ in _lib.Report._fields_}

    def decode_from_queue(self, queue: "FrameQueue", channel: int, channel_param: float, inputs, syndromes, results,
                          iter_counts: np.ndarray | None = None, num_iter_max: int = 100, num_iter_check_parity: int = 10,
                          layout=LAYOUT_INTERLEAVED, pool_frames: int = 0, flags: int = 0, frame_stride: int = 0,
                          lookahead: int = 0):
        """Decode the frames this decoder manages to reserve from a shared queue (one decoder per GPU; call it
        concurrently from one thread or process per decoder).  inputs / syndromes / results / iter_counts are sized
        for the whole queue (or a pool); returns (iter_counts, report) with report['frames_decoded']."""
        n = queue.size
        if isinstance(inputs, np.ndarray):
            in_dtype = np.uint32 if layout == LAYOUT_PACKED_BITS else (np.float16 if flags & FLAG_HALF_INPUT else self.np_dtype)
            inputs = np.ascontiguousarray(inputs, dtype=in_dtype)
        if isinstance(syndromes, np.ndarray):
            syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
        if iter_counts is None:
            iter_counts = np.zeros(n, dtype=np.uint32)
        prm = _lib.DecodeParams(channel=channel, channel_param=channel_param, num_iter_max=num_iter_max,
                                num_iter_check_parity=num_iter_check_parity, input_layout=layout,
                                frame_stride=frame_stride, input_pool_frames=pool_frames, flags=flags,
                                queue_lookahead_frames=lookahead)
        rep = _lib.Report()
        _wait_for_torch(inputs, syndromes, results)
        _check(_lib.lib().ldpc_b200_decode_from_queue(self._h, C.byref(prm), queue._h, _ptr(inputs), _ptr(syndromes),
                                                      _ptr(results), C.byref(rep), _ptr(iter_counts)))
        return iter_counts, {k: getattr(rep, k) for k, _ in _lib.Report._fields_}

    def synthesize_frames(self, channel: int, noise: float, start_idx: int, n_frames: int, d_input, d_syndromes,
                          d_ref_bits):
        """Fill CUDA tensors with the reference harness' frames start_idx.. (generated on the device)."""
        _wait_for_torch(d_input, d_syndromes, d_ref_bits)
        _check(_lib.lib().ldpc_b200_synthesize_frames(self._h, channel, noise, start_idx, n_frames, _ptr(d_input),
                                                      _ptr(d_syndromes), _ptr(d_ref_bits)))

    # ---- kernel-level entry points (parity tests / profiling) ----
    def refill(self, channel, channel_param, staged, syndromes):
        staged = np.ascontiguousarray(staged, dtype=self.np_dtype)
        syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
        _check(_lib.lib().ldpc_b200_debug_refill(self._h, channel, channel_param, staged.shape[1], _ptr(staged),
                                                 _ptr(syndromes)))

    def iterate(self, n_iters: int, final_bits: bool = False, timed: bool = False):
        cn, vn = C.c_float(0), C.c_float(0)
        _check(_lib.lib().ldpc_b200_debug_iterate(self._h, n_iters, int(final_bits),
                                                  C.byref(cn) if timed else None, C.byref(vn) if timed else None))
        return cn.value, vn.value

    def cn_pass(self):
        _check(_lib.lib().ldpc_b200_debug_cn_pass(self._h))

    def vn_pass(self, final_bits: bool = False):
        _check(_lib.lib().ldpc_b200_debug_vn_pass(self._h, int(final_bits)))

    def check(self) -> np.ndarray:
        P = max(32, self.parallel_factor())
        v = np.zeros(P, dtype=np.uint8)
        _check(_lib.lib().ldpc_b200_debug_check(self._h, _ptr(v)))
        return v

    def messages(self, slot: int) -> np.ndarray:
        out = np.zeros(self.code.n_edges, dtype=np.float32)
        _check(_lib.lib().ldpc_b200_debug_read_messages(self._h, slot, _ptr(out)))
        return out

    def priors(self, slot: int) -> np.ndarray:
        out = np.zeros(self.code.n_vars, dtype=np.float32)
        _check(_lib.lib().ldpc_b200_debug_read_priors(self._h, slot, _ptr(out)))
        return out

    def hard(self, slot: int) -> np.ndarray:
        out = np.zeros(self.code.n_vars // 32, dtype=np.uint32)
        _check(_lib.lib().ldpc_b200_debug_read_hard(self._h, slot, _ptr(out)))
        return out

    def phi_table_fp16(self) -> np.ndarray:
        t = np.zeros(32768, dtype=np.uint16)
        _check(_lib.lib().ldpc_b200_phi_table_fp16(self._h, _ptr(t)))
        return t

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.lib().ldpc_b200_decoder_destroy(self._h)
            self._h = None


class FrameQueue:
    """A queue of frames shared by several decoders (include/ldpc_b200.h, "one frame queue, several GPUs"): an atomic
    cursor handing out blocks of consecutive frames.  shm_name: POSIX shared-memory name for one process per GPU (one
    process creates, the others open with create=False); None: private to this process."""

    def __init__(self, n_frames: int = 0, block_frames: int = 0, shm_name: str | None = None, create: bool = True):
        h = C.c_void_p()
        _check(_lib.lib().ldpc_b200_queue_create(shm_name.encode() if shm_name else None, int(create), n_frames,
                                                 block_frames, C.byref(h)))
        self._h = h
        self.size = int(_lib.lib().ldpc_b200_queue_size(h))

    def reset(self):
        _lib.lib().ldpc_b200_queue_reset(self._h)

    def take(self):
        """Reserve the next block by hand: (first_frame, count), count == 0 when the queue is exhausted."""
        first = C.c_uint32(0)
        n = _lib.lib().ldpc_b200_queue_take(self._h, C.byref(first))
        return int(first.value), int(n)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.lib().ldpc_b200_queue_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def decode_multi(decoders, channel: int, channel_param: float, inputs, syndromes, n_frames: int | None = None,
                 num_iter_max: int = 100, num_iter_check_parity: int = 10, results=None, layout=LAYOUT_INTERLEAVED,
                 pool_frames: int = 0, flags: int = 0, frame_stride: int = 0, block_frames: int = 0, lookahead: int = 0):
    """One queue of frames decoded by several decoders (one per GPU) of ONE process: a host thread per decoder and a
    shared atomic cursor inside the library (ldpc_b200_decode_multi).  Host buffers, or buffers every device can
    address.  Returns (results, iter_counts, merged report, per-decoder reports)."""
    d0 = decoders[0]
    if isinstance(inputs, np.ndarray):
        in_dtype = np.uint32 if layout == LAYOUT_PACKED_BITS else (np.float16 if flags & FLAG_HALF_INPUT else d0.np_dtype)
        inputs = np.ascontiguousarray(inputs, dtype=in_dtype)
        if n_frames is None:
            n_frames = inputs.shape[1] if layout == LAYOUT_INTERLEAVED else inputs.shape[0]
    if isinstance(syndromes, np.ndarray):
        syndromes = np.ascontiguousarray(syndromes, dtype=np.uint32)
    assert n_frames is not None
    if results is None:
        results = np.zeros((n_frames, d0.code.n_vars // 32), dtype=np.uint32)
    iters = np.zeros(n_frames, dtype=np.uint32)
    prm = _lib.DecodeParams(channel=channel, channel_param=channel_param, num_iter_max=num_iter_max,
                            num_iter_check_parity=num_iter_check_parity, input_layout=layout, frame_stride=frame_stride,
                            input_pool_frames=pool_frames, flags=flags, queue_lookahead_frames=lookahead)
    handles = (C.c_void_p * len(decoders))(*[d._h for d in decoders])
    rep = _lib.Report()
    per = (_lib.Report * len(decoders))()
    _wait_for_torch(inputs, syndromes, results)
    _check(_lib.lib().ldpc_b200_decode_multi(handles, len(decoders), C.byref(prm), n_frames, _ptr(inputs), _ptr(syndromes),
                                             _ptr(results), C.byref(rep), per, _ptr(iters), block_frames))
    as_dict = lambda r: {k: getattr(r, k) for k, _ in _lib.Report._fields_}  # noqa: E731
    return results, iters, as_dict(rep), [as_dict(r) for r in per]


def phi_abs_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros_like(x)
    _check(_lib.lib().ldpc_b200_debug_phi_abs_f32(_ptr(x), _ptr(y), x.size))
    return y
