"""ctypes binding of libb200rec.so (include/b200rec.h).  No torch types cross this boundary: tensors are passed as
`tensor.data_ptr()` and the current CUDA stream as a raw pointer.

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200REC_LIB") or os.path.join(_HERE, "libb200rec.so")  # B200REC_LIB: A/B builds of the same ABI
_lib = None


class B200RecError(RuntimeError):
    pass


class CsrStruct(C.Structure):
    """mirror of `b200rec_csr` (include/b200rec.h)"""
    _fields_ = [
        ("n_rows", C.c_int32), ("n_cols", C.c_int32), ("nnz", C.c_int32),
        ("rowptr", C.c_void_p), ("colidx", C.c_void_p), ("col_hint", C.c_int32), ("vals", C.c_void_p),
        ("nbr_scale", C.c_void_p), ("row_scale", C.c_void_p), ("eid", C.c_void_p),
        ("n_items", C.c_int32),
        ("item_start", C.c_void_p), ("item_end", C.c_void_p), ("item_dst", C.c_void_p), ("item_row", C.c_void_p),
        ("n_long", C.c_int32),
        ("long_row", C.c_void_p), ("long_slot0", C.c_void_p), ("long_nslot", C.c_void_p), ("long_cnt", C.c_void_p),
        ("n_slots", C.c_int32), ("slot_long", C.c_void_p),
        ("partial", C.c_void_p),
        ("n_passes", C.c_int32), ("pass_ptr", C.c_void_p),
        ("records", C.c_void_p), ("win_start", C.c_void_p), ("pass_win_ptr", C.c_void_p), ("win_counter", C.c_void_p),
    ]


_P, _I32, _I64, _F, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64
_CSRP = C.POINTER(CsrStruct)

# name -> (restype, argtypes); must list every symbol declared in include/b200rec.h
PROTOTYPES = {
    "b200rec_last_error": (C.c_char_p, []),
    "b200rec_version": (C.c_int, []),
    "b200rec_launch_count": (C.c_uint64, []),
    "b200rec_csr_build": (C.c_int, [_P, _P, _I64, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "b200rec_adj_build": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "b200rec_plan_build": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200rec_stream_build": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "b200rec_l2_persist": (C.c_int, [_I64, _P]),
    "b200rec_adj_normalize": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P]),
    "b200rec_spmm_f32": (C.c_int, [_CSRP, _P, _I32, _P, _F, _P, _P, _P, _F, _P]),
    "b200rec_spmm_f32_ex": (C.c_int, [_CSRP, _P, _I32, _P, _F, _P, _P, _P, _F, _P, _P, _P]),
    "b200rec_spmm_f32_sel": (C.c_int, [_CSRP, _P, _I32, _F, _P, _P, _P, _F, _P, _P, _P, _I32, _P]),
    "b200rec_spmm_f32_blocked": (C.c_int, [_CSRP, _P, _I32, _I32, _F, _P, _P, _P, _F, _P, _P]),
    "b200rec_live_items": (C.c_int, [_CSRP, _P, _P, _P, _P]),
    "b200rec_spmm_f32_live": (C.c_int, [_CSRP, _P, _I32, _F, _P, _P, _P, _F, _P, _P, _I32, _P]),
    "b200rec_propagate_fwd": (C.c_int, [_CSRP, _P, _I32, _I32, _P, _P, _P, _P, _P]),
    "b200rec_propagate_bwd": (C.c_int, [_CSRP, _P, _I32, _I32, _P, _P, _P, _P, _P]),
    "b200rec_bpr_sample": (C.c_int, [_P, _P, _I32, _I32, _U64, _P, _I32, _P, _P]),
    "b200rec_mark_rows": (C.c_int, [_P, _I32, _I64, _P, _P]),
    "b200rec_mark_reach": (C.c_int, [_P, _I32, _I64, _P, _P, _P, _P]),
    "b200rec_gather_rows": (C.c_int, [_P, _I32, _P, _I64, _I32, _P, _P, _P]),
    "b200rec_scatter_add_rows": (C.c_int, [_P, _I32, _P, _I64, _I32, _P, _P]),
    "b200rec_bpr_scratch_floats": (C.c_int64, [_I32, _I32]),
    "b200rec_bpr_fwd_bwd": (C.c_int, [_P, _I32, _P, _I32, _I64, _F, _I32, _P, _F, _P, _P, _P, _P, _P]),
    "b200rec_bpr_fwd_bwd_sharded": (C.c_int, [_P, _I32, _P, _I32, _I64, _F, _I32, _P, _F, _P, _P, _P, _P, _P, _I32, _F, _P]),
    "b200rec_bpr_group_rows": (C.c_int, [_P, _I32, _I64, _P, _P]),
    "b200rec_bpr_fwd_bwd_ordered": (C.c_int, [_P, _I32, _P, _I32, _I64, _F, _I32, _P, _F, _P, _P, _P, _P, _P, _I32, _F, _P, _P,
                                              _I32, _P, _F, _P]),
    "b200rec_bpr_l2_emb0_ordered": (C.c_int, [_P, _I32, _P, _I32, _I64, _F, _P, _P, _P, _P, _P, _P]),
    "b200rec_clear_rows": (C.c_int, [_P, _I32, _I64, _P, _I32, _P]),
    "b200rec_bpr_l2_emb0": (C.c_int, [_P, _I32, _P, _I32, _I64, _F, _P, _P, _P, _P]),
    "b200rec_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P]),
    "b200rec_step_advance": (C.c_int, [_P, _P, _P, _P, _I32, _P]),
    "b200rec_dropout_mask": (C.c_int, [_I32, _F, _U64, _P, _P, _P]),
    "b200rec_infonce_workspace_floats": (C.c_int64, [_I32, _I32]),
    "b200rec_infonce_fwd_bwd": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _F, _F, _P, _P, _P, _P, _P]),
    "b200rec_score_dense_f32": (C.c_int, [_P, _P, _I32, _P, _I32, _I32, _P, _P]),
    "b200rec_score_topk_workspace": (C.c_int64, [_I32, _I32, _I32, _I32, _I32]),
    "b200rec_score_topk": (C.c_int, [_P, _P, _I32, _P, _I32, _I32, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "b200rec_spmm_f32_peer": (C.c_int, [_CSRP, _P, _I32, _P, _F, _P, _P, _P, _F, _P, _P, _I32, _P, _P, _P, _I32, _P]),
    "b200rec_adam_step_peer": (C.c_int, [_P, _P, _P, _P, _I64, C.c_double, C.c_double, C.c_double, C.c_double, _P, _I32, _P, _P]),
    "b200rec_peer_alloc": (C.c_int, [_I64, _P, _P]),
    "b200rec_peer_open": (C.c_int, [_P, _P]),
    "b200rec_peer_close": (C.c_int, [_P]),
    "b200rec_peer_free": (C.c_int, [_P]),
    "b200rec_peer_signal": (C.c_int, [_P, _I32, _I32, _P, _I64, _P]),
    "b200rec_peer_wait": (C.c_int, [_P, _I32, _P, _I64, _P]),
    "b200rec_topk_global": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "b200rec_rank_metrics": (C.c_int, [_P, _I32, _I32, _I64, _P, _P, _P, _I32, _P, _P]),
    "b200rec_hit_matrix": (C.c_int, [_P, _I32, _I32, _I64, _P, _P, _P, _P]),
}


def load():
    """dlopen libb200rec.so and bind every prototype; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200RecError(
            "libb200rec.so not found at %s -- build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device (or host) pointer of a tensor / None"""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def check(rc, what=""):
    if rc != 0:
        msg = load().b200rec_last_error().decode("utf-8", "replace")
        raise B200RecError("%s failed (%d): %s" % (what or "b200rec call", rc, msg))


def require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise B200RecError("b200rec needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B200RecError("expected a CUDA tensor, got %s" % t.device)
        if t is not None and not t.is_contiguous():
            raise B200RecError("expected a contiguous tensor")


def launch_count():
    return int(load().b200rec_launch_count())
