"""ctypes binding of libqst.so (the C ABI declared in include/qst.h).

There is no CPU or PyTorch fallback: if the library cannot be loaded every product entry
point raises ``QstLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqst.so")

QST_F32, QST_F16, QST_BF16 = 0, 1, 2
QST_RED_NONE, QST_RED_SUM, QST_RED_MEAN = 0, 1, 2
QST_SCORE_COS, QST_SCORE_DOT, QST_SCORE_EUCLID = 0, 1, 2
QST_PREP_RAW, QST_PREP_COS, QST_PREP_EUCLID_CORPUS, QST_PREP_EUCLID_QUERY = 0, 1, 2, 3
QST_QUAD_SAVED_PER_ROW = 8

REDUCTION_CODES = {"none": QST_RED_NONE, "sum": QST_RED_SUM, "mean": QST_RED_MEAN}


class QstLibraryError(RuntimeError):
    pass


class QstError(RuntimeError):
    pass


class QuadParams(C.Structure):
    _fields_ = [("gamma", C.c_float), ("one_minus_gamma", C.c_float), ("margin_pos_neg", C.c_float),
                ("margin_pos_part", C.c_float), ("margin_part_neg", C.c_float), ("p", C.c_float),
                ("eps", C.c_float), ("swap", C.c_int32)]


class TopkPlan(C.Structure):
    _fields_ = [("Q", C.c_int64), ("N", C.c_int64), ("D", C.c_int64), ("D_pad", C.c_int64),
                ("k", C.c_int32), ("kprime", C.c_int32), ("kunit", C.c_int32), ("cap", C.c_int32),
                ("m_tiles", C.c_int32), ("n_tiles", C.c_int32), ("stripes", C.c_int32),
                ("tiles_per_stripe", C.c_int32), ("units", C.c_int32),
                ("grid", C.c_int32), ("score", C.c_int32),
                ("ctas", C.c_int32), ("rows_per_unit", C.c_int32),
                ("ws_bytes", C.c_size_t), ("off_thr", C.c_size_t), ("off_cnt", C.c_size_t), ("off_uthr", C.c_size_t),
                ("off_cand", C.c_size_t), ("qs", C.c_int32), ("reserved_", C.c_int32)]


QST_MAX_WORLD = 16


class Scatter(C.Structure):
    """``qst_scatter`` of include/qst.h: peer-mapped receive buffers of the ranks of a node."""
    _fields_ = [("base", C.c_void_p * QST_MAX_WORLD), ("world", C.c_int32), ("rank", C.c_int32),
                ("rows_per_block", C.c_int64)]


_P = C.c_void_p
_I64 = C.c_int64
_INT = C.c_int

# every symbol include/qst.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "qst_version": (_INT, []),
    "qst_last_error": (C.c_char_p, []),
    "qst_launch_count": (C.c_longlong, []),
    "qst_device_info": (_INT, [C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT)]),
    "qst_quadruplet_workspace_bytes": (C.c_size_t, []),
    "qst_quadruplet_fwd": (_INT, [_P, _P, _P, _P, _INT, _I64, _I64, C.POINTER(QuadParams), _INT, _P, _P, _P, _P]),
    "qst_quadruplet_bwd": (_INT, [_P, _P, _P, _P, _INT, _I64, _I64, C.POINTER(QuadParams), _INT, _P, _P,
                                   _P, _P, _P, _P, _P]),
    "qst_quadruplet_fwd_bwd": (_INT, [_P, _P, _P, _P, _INT, _I64, _I64, C.POINTER(QuadParams), _INT, C.c_float,
                                       _P, _P, _P, _P, _P, _P, _P]),
    "qst_quadruplet_eval": (_INT, [_P, _P, _P, _P, _INT, _I64, _I64, _P, _P, _P]),
    "qst_padded_dim": (_I64, [_I64]),
    "qst_padded_dim_for": (_I64, [_I64, _INT]),
    "qst_prep_rows": (_INT, [_P, _INT, _I64, _I64, _INT, _P, _P, _P, _P, _P, _P]),
    "qst_topk_plan_make": (_INT, [_I64, _I64, _I64, _INT, _INT, _INT, _INT, C.POINTER(TopkPlan)]),
    "qst_topk_plan_set_kunit": (_INT, [C.POINTER(TopkPlan), _INT]),
    "qst_score_select": (_INT, [C.POINTER(TopkPlan), _P, _P, _P, _P]),
    "qst_score_dense": (_INT, [_P, _I64, _P, _I64, _I64, _P, _P]),
    "qst_dense_scores": (_INT, [_I64, _I64, _I64, _INT, _P, _P, _P, _P, _P, _P]),
    "qst_debug_read_trace": (_INT, [_P, _P, _P, _INT]),
    "qst_score_select_peers": (_INT, [C.POINTER(TopkPlan), _P, _P, _P, _P, C.POINTER(_P), _INT, _P]),
    "qst_peer_buffer_create": (_INT, [C.c_size_t, C.POINTER(_P), C.c_char_p]),
    "qst_peer_buffer_open": (_INT, [C.c_char_p, C.POINTER(_P)]),
    "qst_peer_buffer_clear": (_INT, [_P, C.c_size_t, C.c_size_t, _P]),
    "qst_peer_copy": (_INT, [_P, _P, C.c_size_t, _P]),
    "qst_peer_buffer_close": (_INT, [_P]),
    "qst_peer_buffer_destroy": (_INT, [_P]),
    "qst_finalize_topk": (_INT, [C.POINTER(TopkPlan), _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "qst_finalize_topk_adaptive": (_INT, [C.POINTER(TopkPlan), _INT, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "qst_select_candidates": (_INT, [C.POINTER(TopkPlan), _P, _INT, _I64, _P, _P]),
    "qst_finalize_lists_scratch_bytes": (C.c_size_t, [_I64, _INT]),
    "qst_finalize_lists": (_INT, [_I64, _INT, _INT, _INT, _INT, _INT, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, _P]),
    "qst_select_requests": (_INT, [_I64, _INT, _INT, _INT, _I64, _P, _P, _P, _P, _P]),
    "qst_rescore_requests": (_INT, [_I64, _INT, _I64, _INT, _P, _P, _P, _P, _P, _P, _P]),
    "qst_select_candidates_scatter": (_INT, [C.POINTER(TopkPlan), _P, _INT, _I64, C.POINTER(Scatter), _P]),
    "qst_select_requests_scatter": (_INT, [_I64, _INT, _INT, _INT, _I64, _P, _P, _P, _P, C.POINTER(Scatter), _P]),
    "qst_rescore_requests_scatter": (_INT, [_I64, _INT, _I64, _INT, _P, _P, _P, _P, _P, _P, C.POINTER(Scatter), _P]),
    "qst_peer_barrier": (_INT, [C.POINTER(Scatter), C.c_uint32, _P]),
    "qst_finalize_exact": (_INT, [_I64, _INT, _INT, _INT, _INT, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "qst_exact_rescan_lists": (_INT, [_I64, _I64, _I64, _INT, _INT, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "qst_exact_rescan_workspace_bytes": (C.c_size_t, [_I64, _INT]),
    "qst_exact_rescan": (_INT, [_I64, _I64, _I64, _INT, _INT, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "qst_merge_topk": (_INT, [_P, _P, _INT, _I64, _INT, _P, _P, _P]),
    "qst_comm_available": (_INT, []),
    "qst_comm_nccl_version": (_INT, []),
    "qst_comm_unique_id": (_INT, [C.c_char_p]),
    "qst_comm_init": (_INT, [C.c_char_p, _INT, _INT, C.POINTER(_P)]),
    "qst_comm_destroy": (_INT, [_P]),
    "qst_comm_world": (_INT, [_P]),
    "qst_comm_rank": (_INT, [_P]),
    "qst_comm_allgather": (_INT, [_P, _P, _P, C.c_size_t, _P]),
    "qst_comm_alltoall": (_INT, [_P, _P, _P, C.c_size_t, _P]),
    "qst_comm_allreduce_max_f32": (_INT, [_P, _P, _P, C.c_size_t, _P]),
    "qst_allgather_topk": (_INT, [_P, _P, _P, _I64, _INT, _P, _P, _P]),
    "qst_exchange_candidates": (_INT, [_P, _P, _P, _I64, _INT, _P]),
    "qst_ir_metrics": (_INT, [_P, _I64, _INT, _P, _P, _P, _INT, _P, _P, _P, _P]),
}

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Load libqst.so (once).  Raises QstLibraryError when it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            if build_if_missing:
                from . import build as _build
                _build.build()
            else:
                raise QstLibraryError(
                    f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                    "(there is no CPU fallback for this path)")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:
            raise QstLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise QstLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().qst_last_error()
        raise QstError(f"libqst error {rc}: {msg.decode(errors='replace') if msg else '?'}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL; an int is taken as a raw device pointer)."""
    if t is None or isinstance(t, int):
        return t
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(dtype) -> int:
    import torch
    if dtype == torch.float32:
        return QST_F32
    if dtype == torch.float16:
        return QST_F16
    if dtype == torch.bfloat16:
        return QST_BF16
    raise TypeError(f"unsupported dtype {dtype} (float32, float16, bfloat16 only)")


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise QstError("this path runs on a CUDA device only: got a CPU tensor "
                           "(no CPU fallback; move the tensor to the GPU)")
