"""Device-resident sparse operands of the hot path and their construction.

  * `CsrOperand`       -- int32 CSR + the load-balanced work decomposition libb200rec's SpMM consumes.
  * `build_norm_adj`   -- D^-1/2 A D^-1/2 of the user-item graph (utils.py:42-50 generate_daj_mat +
                          model.py:89-98 LightGCN.generate_graph): rows [users | items], nnz = 2E, bit-wise symmetric.
  * `build_feat`       -- IGCN's template feature matrix F and its transpose (model.py:4139-4175 generate_feat).

HBM layout: rowptr int32[N+1], colidx int32[nnz] ascending inside a row (the order of the reference's coalesced COO,
utils.py:33-39), vals fp32[nnz]; 8 B/edge instead of the reference's 20 B/edge int64 COO.  Index sorting uses torch
device ops (one-time setup plumbing); the edge values come from the library (b200rec_adj_normalize).
"""
import ctypes as C

import numpy as np
import torch

from . import _abi

CHUNK_MAX = 1024


def auto_chunk(nnz, d=None):
    """Rows longer than `chunk` are split into chunk-sized work items (deterministic ordered reduce by the lane group that
    parks the row's last chunk).  One lane group walks its item serially, about 8 neighbour rows per memory round trip, so
    the longest item is a critical path (ncu, C2 shape with 1024-edge items: SMs 54 % active); every extra chunk costs a
    partial-row round trip through L2.  Measured sweet spot: a few items per resident lane group, within [64, 1024]."""
    import os
    if os.environ.get("B200REC_CHUNK"):
        return int(os.environ["B200REC_CHUNK"])
    target = max(1, nnz // (148 * 32 * 4))
    c = 64
    while c < target and c < CHUNK_MAX:
        c *= 2
    return c


class CsrOperand:
    def __init__(self, rowptr, colidx, n_cols, vals=None, nbr_scale=None, row_scale=None, eid=None, chunk=None,
                 max_d=256, phase_split=0, d=None):
        _abi.require_cuda(rowptr, colidx, vals, nbr_scale, row_scale, eid)
        assert rowptr.dtype == torch.int32 and colidx.dtype == torch.int32
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.nbr_scale, self.row_scale, self.eid = nbr_scale, row_scale, eid
        self.n_rows = rowptr.numel() - 1
        self.n_cols = int(n_cols)
        self.nnz = colidx.numel()
        self.device = rowptr.device
        self.chunk = chunk if chunk is not None else auto_chunk(self.nnz, d)
        self.phase_split = int(phase_split)
        self.col_hint = 0
        self._build_plan(max_d)
        self._struct = None

    def _build_plan(self, max_d):
        lib = _abi.load()
        rp = self.rowptr.cpu().numpy()
        n_items, n_long, n_slots = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        null = C.c_void_p(0)
        _abi.check(lib.b200rec_plan_build_host(rp.ctypes.data, self.n_rows, self.chunk, self.phase_split,
                                               C.addressof(n_items),
                                               C.addressof(n_long), C.addressof(n_slots), null, null, null, null, null,
                                               null, null, null), "plan_build_host(size)")
        ni, nl, ns = n_items.value, n_long.value, n_slots.value
        a = [np.empty(max(ni, 1), dtype=np.int32) for _ in range(4)]
        b = [np.empty(max(nl, 1), dtype=np.int32) for _ in range(3)]
        sl = np.zeros(max(ns, 1), dtype=np.int32)
        _abi.check(lib.b200rec_plan_build_host(rp.ctypes.data, self.n_rows, self.chunk, self.phase_split,
                                               C.addressof(n_items),
                                               C.addressof(n_long), C.addressof(n_slots),
                                               a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data,
                                               b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data, sl.ctypes.data),
                   "plan_build_host")
        dev = self.device
        self.n_items, self.n_long, self.n_slots = ni, nl, ns
        self.item_start, self.item_end, self.item_dst, self.item_row = (torch.from_numpy(x[:max(ni, 1)]).to(dev) for x in a)
        self.long_row, self.long_slot0, self.long_nslot = (torch.from_numpy(x[:max(nl, 1)]).to(dev) for x in b)
        self.slot_long = torch.from_numpy(sl).to(dev)
        self.long_cnt = torch.zeros(max(nl, 1), dtype=torch.int32, device=dev)
        self.partial = torch.empty((max(ns, 1), max_d), dtype=torch.float32, device=dev) if ns else None
        self.max_d = max_d

    def struct(self):
        if self._struct is None:
            s = _abi.CsrStruct()
            s.n_rows, s.n_cols, s.nnz = self.n_rows, self.n_cols, self.nnz
            p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
            s.rowptr, s.vals = p(self.rowptr), p(self.vals)
            s.colidx = p(self.colidx_enc if self.col_hint else self.colidx)
            s.col_hint = self.col_hint
            s.nbr_scale, s.row_scale, s.eid = p(self.nbr_scale), p(self.row_scale), p(self.eid)
            s.n_items = self.n_items
            s.item_start, s.item_end, s.item_dst = p(self.item_start), p(self.item_end), p(self.item_dst)
            s.item_row = p(self.item_row)
            s.n_long = self.n_long
            s.long_row, s.long_slot0, s.long_nslot = p(self.long_row), p(self.long_slot0), p(self.long_nslot)
            s.n_slots = self.n_slots
            s.slot_long, s.long_cnt = p(self.slot_long), p(self.long_cnt)
            s.partial = p(self.partial)
            self._struct = s
        return self._struct

    def apply_cache_hints(self, d, n_users=None):
        """Flag the hottest gathered rows (bit 31 of a private copy of colidx) for the kernel's cache-policy loads.
        Table larger than L2 -> hint 2: the top rows that fit ~half of L2 are kept with L2::evict_last while every other
        row streams through evict_first.  Table L2-resident -> hint 1: the top rows that fit L1 allocate there, the
        rest bypass L1.  Users and items are ranked separately: a row gathers only from the other side, and the
        plan runs the two sides as two phases."""
        if self.n_rows != self.n_cols or self.nnz == 0:
            return self
        l2 = torch.cuda.get_device_properties(self.device).L2_cache_size
        row_bytes = d * 4
        import os
        frac = float(os.environ.get("B200REC_HOT_FRAC", "0"))  # measured on C4: hints 11.8-12.1 ms/layer vs 11.6 without -> off
        if self.n_cols * row_bytes > 0.6 * l2 and frac > 0:
            hint, k = 2, int(frac * l2 / row_bytes)
        elif os.environ.get("B200REC_L1_HINT", "0") == "1":
            hint, k = 1, int(160 * 1024 / row_bytes)
        else:
            # L2-resident table: measured on the C2 shape, bypassing L1 for cold rows is a loss (0.372 -> 0.507 ms/step)
            self.col_hint, self._struct = 0, None
            return self
        deg = (self.rowptr[1:] - self.rowptr[:-1]).long()
        hot = torch.zeros(self.n_cols, dtype=torch.bool, device=self.device)
        sides = [(0, self.n_cols)] if not n_users else [(0, n_users), (n_users, self.n_cols)]
        for lo, hi in sides:
            kk = min(k, hi - lo)
            if kk > 0:
                hot[lo + torch.topk(deg[lo:hi], kk).indices] = True
        enc = self.colidx.clone()
        enc[hot[self.colidx.long()]] |= -2 ** 31
        self.colidx_enc, self.col_hint, self.hot_rows, self._struct = enc, hint, int(hot.sum()), None
        return self

    def replan(self, chunk):
        """the same matrix with a different hub-row chunk (own work plan and scratch, shared CSR arrays)"""
        o = CsrOperand(self.rowptr, self.colidx, self.n_cols, vals=self.vals, nbr_scale=self.nbr_scale,
                       row_scale=self.row_scale, eid=self.eid, chunk=chunk, max_d=self.max_d, phase_split=self.phase_split)
        if self.col_hint:
            o.colidx_enc, o.col_hint = self.colidx_enc, self.col_hint
        return o

    def with_scales(self, nbr_scale=None, row_scale=None):
        """same structure/plan, different per-node scale vectors (IGCN anneal: F's values change every epoch)"""
        o = object.__new__(CsrOperand)
        o.__dict__.update(self.__dict__)
        o.nbr_scale, o.row_scale, o._struct = nbr_scale, row_scale, None
        return o

    def row_slice(self, lo, hi=None):
        """rows [lo, hi) -- or a list of such ranges -- as an operand that writes into the FULL output table at global
        row ids (multi-GPU row partition: per-row sums are unchanged, so P-rank results equal 1-rank bit for bit)."""
        ranges = [(lo, hi)] if hi is not None else list(lo)
        o = object.__new__(CsrOperand)
        o.__dict__.update(self.__dict__)
        o._struct = None
        row = self.item_row[:self.n_items]
        keep = torch.zeros_like(row, dtype=torch.bool)
        for a, b in ranges:
            keep |= (row >= a) & (row < b)
        o.item_start = self.item_start[:self.n_items][keep].contiguous()
        o.item_end = self.item_end[:self.n_items][keep].contiguous()
        o.item_dst = self.item_dst[:self.n_items][keep].contiguous()
        o.item_row = self.item_row[:self.n_items][keep].contiguous()
        o.n_items = int(o.item_start.numel())
        if o.n_items == 0:
            for name in ("item_start", "item_end", "item_dst", "item_row"):
                setattr(o, name, torch.zeros(1, dtype=torch.int32, device=self.device))
        return o

    # ---- views for API compatibility / tests (the reference's norm_adj is a coalesced sparse COO tensor) ----
    @property
    def shape(self):
        return torch.Size([self.n_rows, self.n_cols])

    def _nnz(self):
        return self.nnz

    def coalesce(self):
        return self

    def indices(self):
        r, c, _ = self.to_coo()
        return torch.stack([r, c])

    def values(self):
        return self.vals

    def to_coo(self):
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device),
                                       (self.rowptr[1:] - self.rowptr[:-1]).long())
        return rows, self.colidx.long(), self.vals


def _rowptr_from_sorted_rows(rows, n_rows):
    counts = torch.bincount(rows, minlength=n_rows)
    rp = torch.zeros(n_rows + 1, dtype=torch.int64, device=rows.device)
    torch.cumsum(counts, 0, out=rp[1:])
    return rp


def _coalesce(rows, cols, n_rows, n_cols):
    """sort by (row, col), merge duplicates -> (rowptr int32, colidx int32, mult fp32 or None)"""
    key = rows * n_cols + cols
    key, _ = torch.sort(key)
    uniq, counts = torch.unique_consecutive(key, return_counts=True)
    r = torch.div(uniq, n_cols, rounding_mode="floor")
    c = uniq - r * n_cols
    mult = None
    if uniq.numel() != key.numel():
        mult = counts.to(torch.float32)
    rp = _rowptr_from_sorted_rows(r, n_rows)
    assert int(rp[-1]) < 2 ** 31, "nnz must fit int32"
    return rp.to(torch.int32), c.to(torch.int32), mult


def build_norm_adj(n_users, n_items, users, items, device=None, chunk=None, d=None):
    """users/items: int64 tensors of the E train pairs (train_array, dataset.py:149-151), any order, duplicates allowed
    (they are summed, like scipy's COO->CSR in utils.py:47-49).  Returns (CsrOperand with .vals/.dinv, mult)."""
    device = torch.device(device) if device is not None else users.device
    users, items = users.to(device, torch.int64), items.to(device, torch.int64)
    n = n_users + n_items
    rows = torch.cat([users, items + n_users])
    cols = torch.cat([items + n_users, users])
    rowptr, colidx, mult = _coalesce(rows, cols, n, n)
    dinv = torch.empty(n, dtype=torch.float32, device=device)
    vals = torch.empty(colidx.numel(), dtype=torch.float32, device=device)
    lib = _abi.load()
    _abi.require_cuda(rowptr)
    _abi.check(lib.b200rec_adj_normalize(_abi.ptr(rowptr), _abi.ptr(colidx), _abi.ptr(mult), n, _abi.ptr(dinv),
                                         _abi.ptr(vals), _abi.stream_ptr()), "adj_normalize")
    import os
    op = CsrOperand(rowptr, colidx, n, vals=vals, chunk=chunk, d=d,
                    phase_split=0 if os.environ.get("B200REC_NO_PHASE", "0") == "1" else n_users)
    if d is not None:
        op.apply_cache_hints(d, n_users)
    op.dinv = dinv
    op.mult = mult
    return op


class FeatOperand:
    """IGCN's F [N, T] (T = template users + template items + 2) and F^T, structure only: every value of row r is
    row_sum[r] ** ((alpha-1)/2 - 0.5) (model.py:4127-4130), so the kernels take a per-row scale vector instead of
    per-edge values; F^T stores the forward edge position (`eid`) so both directions share one dropout bitmask."""

    def __init__(self, fwd, bwd, row_sum):
        self.fwd, self.bwd, self.row_sum = fwd, bwd, row_sum
        self.n_rows, self.n_cols, self.nnz = fwd.n_rows, fwd.n_cols, fwd.nnz

    def row_scale(self, alpha):
        return torch.pow(self.row_sum, (alpha - 1.) / 2. - 0.5)


def build_feat(n_users, n_items, users, items, user_tmpl, item_tmpl, device=None):
    """user_tmpl / item_tmpl: int64 [n_users] / [n_items], template column of each node or -1 (the reference's
    user_map / item_map dicts).  Entry list follows model.py:4160-4169."""
    device = torch.device(device) if device is not None else users.device
    users, items = users.to(device, torch.int64), items.to(device, torch.int64)
    user_tmpl, item_tmpl = user_tmpl.to(device), item_tmpl.to(device)
    tu, ti = int((user_tmpl >= 0).sum()), int((item_tmpl >= 0).sum())
    t = tu + ti + 2
    n = n_users + n_items
    mi = item_tmpl[items] >= 0
    mu = user_tmpl[users] >= 0
    ar_u = torch.arange(n_users, device=device)
    ar_i = torch.arange(n_items, device=device)
    rows = torch.cat([users[mi], n_users + items[mu], ar_u, n_users + ar_i])
    cols = torch.cat([tu + item_tmpl[items[mi]], user_tmpl[users[mu]],
                      torch.full_like(ar_u, tu + ti), torch.full_like(ar_i, tu + ti + 1)])
    rowptr, colidx, mult = _coalesce(rows, cols, n, t)
    if mult is None:
        row_sum = (rowptr[1:] - rowptr[:-1]).to(torch.float32)
    else:  # np.sum(feat, axis=1) counts duplicates
        r = torch.repeat_interleave(torch.arange(n, device=device), (rowptr[1:] - rowptr[:-1]).long())
        row_sum = torch.zeros(n, dtype=torch.float32, device=device).index_add_(0, r, mult)
    fwd = CsrOperand(rowptr, colidx, t)
    # transpose: sort edges by (col, row); remember the forward position of every edge
    nnz = colidx.numel()
    r = torch.repeat_interleave(torch.arange(n, device=device), (rowptr[1:] - rowptr[:-1]).long())
    key = colidx.long() * n + r
    key, eid = torch.sort(key)
    tc = torch.div(key, n, rounding_mode="floor")
    tr = key - tc * n
    trp = _rowptr_from_sorted_rows(tc, t).to(torch.int32)
    bwd = CsrOperand(trp, tr.to(torch.int32), n, eid=eid.to(torch.int32))
    assert nnz == bwd.nnz
    return FeatOperand(fwd, bwd, row_sum), tu, ti
