"""Loader of the native library (libldpc_b200.so).  There is no Python or CPU fallback:
if the CUDA extension is missing this raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDPC_B200_LIB") or os.path.join(_HERE, "libldpc_b200.so")  # override: A/B kernel experiments
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_INVALID, ERR_CODE, ERR_MULTIPLE_32, ERR_CUDA, ERR_NOMEM, ERR_IO, ERR_ALIST = range(8)
FP32, FP16 = 0, 1
AWGN, BSC, LLR = 0, 1, 2
LAYOUT_INTERLEAVED, LAYOUT_FRAME_MAJOR, LAYOUT_PACKED_BITS = 0, 1, 2
FLAG_SIMD_PAIRS = 1
FLAG_TIME_KERNELS = 2
FLAG_NO_GRAPH = 4
FLAG_EARLY_RETIRE = 8
FLAG_HALF_INPUT = 16
ORDER_NATURAL, ORDER_LOCALITY = 0, 1


class CodeInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_checks", "n_vars", "n_edges", "n_erased_vars", "n_erased_checks",
                                          "max_deg_var", "max_deg_check")] + \
               [(n, C.POINTER(C.c_uint32)) for n in ("row_ptr", "col_ptr", "in_to_out", "out_to_in", "out_edge_to_var")]


class Config(C.Structure):
    _fields_ = [("precision", C.c_int32), ("max_log2_frames", C.c_uint32), ("device", C.c_int32),
                ("mem_headroom_pct", C.c_uint32), ("storage_order", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class DecodeParams(C.Structure):
    _fields_ = [("channel", C.c_int32), ("channel_param", C.c_float), ("num_iter_max", C.c_uint32),
                ("num_iter_check_parity", C.c_uint32), ("input_layout", C.c_int32), ("frame_stride", C.c_uint64),
                ("input_pool_frames", C.c_uint32), ("flags", C.c_uint32), ("queue_lookahead_frames", C.c_uint32),
                ("reserved", C.c_uint32 * 3)]


class Report(C.Structure):
    _fields_ = [("max_iter", C.c_uint32), ("min_iter", C.c_uint32), ("avg_iter", C.c_float),
                ("iter_time_per_vector", C.c_float), ("global_iters", C.c_uint32), ("n_checks_run", C.c_uint32),
                ("elapsed_s", C.c_double), ("device_s", C.c_double), ("cn_s", C.c_double), ("vn_s", C.c_double),
                ("cn_launches", C.c_uint64), ("vn_launches", C.c_uint64), ("other_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("check_s", C.c_double), ("retire_s", C.c_double), ("refill_s", C.c_double),
                ("frames_refilled", C.c_uint64), ("graph_launches", C.c_uint64),
                ("slot_iterations", C.c_uint64), ("cn_refill_launches", C.c_uint64), ("cn_refill_s", C.c_double),
                ("frames_decoded", C.c_uint64), ("host_stage_s", C.c_double), ("first_batch_stage_s", C.c_double)]


# every symbol include/ldpc_b200.h declares: name -> (restype, argtypes)
_VP, _U32P = C.c_void_p, C.c_void_p
SYMBOLS = {
    "ldpc_b200_code_load_alist": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "ldpc_b200_code_parse_alist": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ldpc_b200_code_from_csr": (C.c_int, [C.c_uint32, C.c_uint32, _U32P, _U32P, C.c_uint32, C.POINTER(C.c_void_p)]),
    "ldpc_b200_code_destroy": (None, [C.c_void_p]),
    "ldpc_b200_code_get_info": (C.c_int, [C.c_void_p, C.POINTER(CodeInfo)]),
    "ldpc_b200_code_set_erased_vars": (None, [C.c_void_p, C.c_uint32]),
    "ldpc_b200_code_locality_order": (C.c_int, [C.c_void_p, _U32P, _U32P]),
    "ldpc_b200_decoder_create": (C.c_int, [C.c_void_p, C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "ldpc_b200_decoder_destroy": (None, [C.c_void_p]),
    "ldpc_b200_decoder_frames_in_flight": (C.c_uint32, [C.c_void_p]),
    "ldpc_b200_decoder_set_erased_vars": (None, [C.c_void_p, C.c_uint32]),
    "ldpc_b200_decoder_prepare_host_io": (C.c_int, [C.c_void_p]),
    "ldpc_b200_decode": (C.c_int, [C.c_void_p, C.POINTER(DecodeParams), C.c_uint32, _VP, _VP, _VP,
                                   C.POINTER(Report), _VP]),
    "ldpc_b200_queue_create": (C.c_int, [C.c_char_p, C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "ldpc_b200_queue_destroy": (None, [C.c_void_p]),
    "ldpc_b200_queue_size": (C.c_uint32, [C.c_void_p]),
    "ldpc_b200_queue_reset": (None, [C.c_void_p]),
    "ldpc_b200_queue_take": (C.c_uint32, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "ldpc_b200_decode_from_queue": (C.c_int, [C.c_void_p, C.POINTER(DecodeParams), C.c_void_p, _VP, _VP, _VP,
                                              C.POINTER(Report), _VP]),
    "ldpc_b200_decode_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(DecodeParams), C.c_uint32, _VP, _VP,
                                         _VP, C.POINTER(Report), C.POINTER(Report), _VP, C.c_uint32]),
    "ldpc_b200_debug_refill": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_uint32, _VP, _VP]),
    "ldpc_b200_debug_iterate": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(C.c_float),
                                          C.POINTER(C.c_float)]),
    "ldpc_b200_debug_cn_pass": (C.c_int, [C.c_void_p]),
    "ldpc_b200_debug_vn_pass": (C.c_int, [C.c_void_p, C.c_int32]),
    "ldpc_b200_debug_check": (C.c_int, [C.c_void_p, _VP]),
    "ldpc_b200_debug_read_messages": (C.c_int, [C.c_void_p, C.c_uint32, _VP]),
    "ldpc_b200_debug_read_priors": (C.c_int, [C.c_void_p, C.c_uint32, _VP]),
    "ldpc_b200_debug_read_hard": (C.c_int, [C.c_void_p, C.c_uint32, _VP]),
    "ldpc_b200_phi_table_fp16": (C.c_int, [C.c_void_p, _VP]),
    "ldpc_b200_debug_phi_abs_f32": (C.c_int, [_VP, _VP, C.c_size_t]),
    "ldpc_b200_synthesize_frames": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, _VP, _VP, _VP]),
    "ldpc_b200_bytes_per_frame_iteration": (C.c_uint64, [C.c_void_p]),
    "ldpc_b200_last_error": (C.c_char_p, []),
    "ldpc_b200_version": (C.c_char_p, []),
    "ldpc_b200_abi_sizes": (None, [C.POINTER(C.c_uint32)]),
}


def build(force: bool = False) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", CSRC, "clean"])
    subprocess.check_call(["make", "-s", "-C", CSRC])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
                "The B200 decoder has no CPU or Python fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
