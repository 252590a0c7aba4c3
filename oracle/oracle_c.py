"""ctypes loader for oracle/_build/liboracle.so (oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "oracle_c.c")):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def bpr_sample(user_ptr, user_items, n_users, n_items, seed, step, batch):
    up = np.ascontiguousarray(user_ptr, dtype=np.int32)
    ui = np.ascontiguousarray(user_items, dtype=np.int32)
    out = np.empty((batch, 3), dtype=np.int64)
    lib().oracle_bpr_sample(_p(up), _p(ui), C.c_int32(n_users), C.c_int32(n_items), C.c_uint64(seed), C.c_uint64(step),
                            C.c_int32(batch), _p(out))
    return out


def dropout_mask(nnz, p, seed, step):
    bits = np.empty((nnz + 31) // 32, dtype=np.uint32)
    lib().oracle_dropout_mask(C.c_int32(nnz), C.c_float(p), C.c_uint64(seed), C.c_uint64(step), _p(bits))
    return bits


def unpack_bits(bits, nnz):
    b = np.unpackbits(bits.view(np.uint8), bitorder="little")
    return b[:nnz].astype(bool)


def scores(rep_users, users, rep_items):
    ru = np.ascontiguousarray(rep_users, dtype=np.float32)
    ri = np.ascontiguousarray(rep_items, dtype=np.float32)
    us = np.ascontiguousarray(users, dtype=np.int64)
    out = np.empty((len(us), ri.shape[0]), dtype=np.float32)
    lib().oracle_scores_f32(_p(ru), _p(us), C.c_int32(len(us)), _p(ri), C.c_int32(ri.shape[0]), C.c_int32(ri.shape[1]), _p(out))
    return out


def mask_topk(sc, users, k, excl_a=None, excl_b=None, banned=(0, 0)):
    sc = np.ascontiguousarray(sc, dtype=np.float32)
    us = np.ascontiguousarray(users, dtype=np.int64)
    nb, ni = sc.shape
    ids = np.empty((nb, k), dtype=np.int32)
    vals = np.empty((nb, k), dtype=np.float32)
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in excl_a] if excl_a is not None else [None, None]
    b = [np.ascontiguousarray(x, dtype=np.int32) for x in excl_b] if excl_b is not None else [None, None]
    lib().oracle_mask_topk(_p(sc), _p(us), C.c_int32(nb), C.c_int32(ni), _p(a[0]), _p(a[1]), _p(b[0]), _p(b[1]),
                           C.c_int32(banned[0]), C.c_int32(banned[1]), C.c_int32(k), _p(ids), _p(vals))
    return ids, vals


def spmm_csr(rowptr, colidx, vals, x, chunk=1024):
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    va = np.ascontiguousarray(vals, dtype=np.float32) if vals is not None else None
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty((len(rp) - 1, x.shape[1]), dtype=np.float32)
    lib().oracle_spmm_csr(_p(rp), _p(ci), _p(va), C.c_int32(len(rp) - 1), _p(x), C.c_int32(x.shape[1]), C.c_int32(chunk), _p(y))
    return y


def infonce(q, k, temperature=0.1, grads=True):
    """(loss, dq, dk) of the reference's InfoNCE call (float64), see oracle_c.c"""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    n, d = q.shape
    gq = np.empty((n, d), dtype=np.float64) if grads else None
    gk = np.empty((n, d), dtype=np.float64) if grads else None
    f = lib().oracle_infonce
    f.restype = C.c_double
    loss = f(_p(q), _p(k), C.c_int32(n), C.c_int32(d), C.c_double(temperature), _p(gq), _p(gk))
    return float(loss), gq, gk


def rank_metrics(rec, eval_ptr, eval_idx, topks):
    """{'Precision'|'Recall'|'NDCG': {k: value}} like calculate_metrics (trainer.py:115-144), plus the user count"""
    rec = np.ascontiguousarray(rec, dtype=np.int32)
    ptr = np.ascontiguousarray(eval_ptr, dtype=np.int32)
    idx = np.ascontiguousarray(eval_idx, dtype=np.int32) if len(eval_idx) else np.zeros(1, dtype=np.int32)
    tk = np.ascontiguousarray(list(topks), dtype=np.int32)
    out = np.empty((3, len(tk)), dtype=np.float64)
    f = lib().oracle_rank_metrics
    f.restype = C.c_int32
    n = f(_p(rec), C.c_int32(rec.shape[0]), C.c_int32(rec.shape[1]), _p(ptr), _p(idx), _p(tk), C.c_int32(len(tk)), _p(out))
    res = {name: {int(kk): float(out[r, t]) for t, kk in enumerate(tk)} for r, name in enumerate(("Precision", "Recall", "NDCG"))}
    return res, int(n)
