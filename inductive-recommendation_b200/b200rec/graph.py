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
COLSORT_NNZ = 1 << 24   # b200rec_plan_build orders the hub chunks of graphs at least this large by source column (plan.cu)


def auto_chunk(nnz, d=None):
    """Rows longer than `chunk` are split into chunk-sized work items (deterministic ordered reduce by the lane group that
    parks the row's last chunk).  One lane group walks its item serially, about 8 neighbour rows per memory round trip, so
    the longest item is a critical path (ncu, C2 shape with 1024-edge items: SMs 54 % active); every extra chunk costs a
    partial-row round trip through L2.  Measured sweet spot: a few items per resident lane group, within [64, 1024]
    (C3: 512 edges 0.963 ms/step, 256 0.977).  Graphs of >= 16 M entries: the full chunks of all hub rows are scheduled by
    their first source column (b200rec_plan_build), and the shorter a chunk, the narrower the window of source rows the
    chunks in flight share -- C4 step: 128 / 256 / 512 / 1024 edges = 66.2 / 64.2 / 64.4 / 65.6 ms (256 WITHOUT the column
    order 70.8) -- so the chunk is capped at 256 there."""
    import os
    if os.environ.get("B200REC_CHUNK"):
        return int(os.environ["B200REC_CHUNK"])
    target = max(1, nnz // (148 * 32 * 4))
    cap = 256 if nnz >= COLSORT_NNZ else CHUNK_MAX
    c = 64
    while c < target and c < cap:
        c *= 2
    return c


def reach_gain(item_deg, n_users, batch):
    """Share of the item-side edges that lies OUTSIDE one hop of an average BPR batch (engine: whether the one-hop mask of
    ops.mark_reach pays).  An item of degree d is missed by `batch` uniformly drawn users with probability (1 - d/U)^batch,
    so the share is sum_i d_i (1 - d_i/U)^batch / sum_i d_i.  item_deg: integer tensor [n_items] (any device)."""
    deg = item_deg.double()
    miss = torch.exp(batch * torch.log1p(-(deg / float(n_users)).clamp(max=1.0 - 1e-12)))
    return float((deg * miss).sum() / deg.sum().clamp(min=1.0))


def blocking_policy(n_cols, d):
    """(sweep width, source rows per block) for a gathered table [n_cols, d] fp32, or None when the table is left to the
    cache as it is -- the default.

    A table much larger than L2 (C4: 3M x 128 x 4 B = 1.5 GB against 126 MB) is gathered at ~24 % L2 hit rate and the
    single-pass SpMM moves ~16x its algorithmic bytes over HBM.  Cutting the source rows into L2-sized blocks (one pass
    each, b200rec_spmm_f32_blocked; wide tables optionally swept in column slices of `sweep` floats) turns those gathers
    into L2 hits -- random 512 B rows come out of a <= 80 MB table at 17.5 TB/s against 7.5 TB/s out of 1.5 GB
    (profiles/r02_l2_gather_ceiling.txt) -- but every pass cuts the rows into ~5-edge segments, and three builds of the
    blocked kernel (profiles/r02_spmm_block_sweep.txt) all lost to the single-pass kernel on C4, ms per layer:
        D=128  single pass 11.1-11.9 | one work item per segment 12.5-34 | record stream, static windows 22.8-24.2 |
               record stream, dynamic windows + DRAM-row prefetch 17.9-19.7
        D=16   single pass 1.86      | one work item per segment 1.76    | record stream 8.8 -> 3.6
    The single-pass kernel already runs at the HBM ceiling (81 % of peak DRAM throughput, and FASTER than a pure random
    gather of the same footprint thanks to the power-law reuse), so the blocked form has to beat it on issue slots and
    latency as well, and with one carried-sum round trip per ~5 edges it does not.  Blocking therefore stays opt-in
    (B200REC_BLOCK_MB=<MB per block>, B200REC_BLOCK_MAX_D, B200REC_SWEEP_D) for experiments and is exercised by the tests."""
    import os
    block_mb = float(os.environ.get("B200REC_BLOCK_MB", "0"))
    if block_mb <= 0:
        return None
    sweep = int(os.environ.get("B200REC_SWEEP_D", "0")) or d
    sweep = min(d, max(8, sweep))
    if d % sweep:
        sweep = d
    if sweep > int(os.environ.get("B200REC_BLOCK_MAX_D", "256")):
        return None
    block_bytes = block_mb * (1 << 20)
    if n_cols * sweep * 4 <= 1.5 * block_bytes:
        return None
    return sweep, max(1024, int(block_bytes // (sweep * 4)))


class CsrOperand:
    def __init__(self, rowptr, colidx, n_cols, vals=None, nbr_scale=None, row_scale=None, eid=None, chunk=None,
                 max_d=256, phase_split=0, d=None, col_bounds=None, row_order=None, window=None):
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
        self.col_bounds = None if col_bounds is None else np.ascontiguousarray(col_bounds, dtype=np.int32)
        # a column-blocked operand (col_bounds) is executed from its record stream: its plan is built in row order
        self.row_order = int(col_bounds is not None) if row_order is None else int(row_order)
        self.records = self.win_start = self.pass_win_ptr = self.win_counter = None
        self.col_hint, self.colidx_enc, self.hot_rows = 0, None, 0
        self._build_plan(max_d)
        if self.col_bounds is not None and self.vals is not None and self.row_order:
            import os
            self._build_stream(int(window if window is not None else os.environ.get("B200REC_STREAM_WINDOW", "128")))
        self._struct = None
        self._blocked = {}   # sweep width -> column-blocked twin of this operand (shared CSR arrays, own plan)
        self._carry = {}     # row stride -> carried-sum scratch [n_rows, ld]

    def _build_plan(self, max_d):
        """b200rec_plan_build on the device: sizes first, then the arrays"""
        lib = _abi.load()
        dev = self.device
        nb = 1 if self.col_bounds is None else len(self.col_bounds) - 1
        bounds = C.c_void_p(0) if self.col_bounds is None else C.c_void_p(self.col_bounds.ctypes.data)
        sizes = (C.c_int32 * 3)()
        null = C.c_void_p(0)
        with torch.cuda.device(dev):
            _abi.check(lib.b200rec_plan_build(_abi.ptr(self.rowptr), _abi.ptr(self.colidx), self.n_rows, self.chunk,
                                              self.phase_split, bounds, nb, self.row_order, sizes, null, null, null, null,
                                              null, null, null, null, null, _abi.stream_ptr()), "plan_build(size)")
            ni, nl, ns = sizes[0], sizes[1], sizes[2]
            i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=dev)  # noqa: E731
            self.item_start, self.item_end, self.item_dst, self.item_row = i32(ni), i32(ni), i32(ni), i32(ni)
            self.long_row, self.long_slot0, self.long_nslot, self.slot_long = i32(nl), i32(nl), i32(nl), i32(ns)
            self.pass_ptr = np.zeros(nb + 1, dtype=np.int32)
            _abi.check(lib.b200rec_plan_build(_abi.ptr(self.rowptr), _abi.ptr(self.colidx), self.n_rows, self.chunk,
                                              self.phase_split, bounds, nb, self.row_order, sizes, _abi.ptr(self.item_start),
                                              _abi.ptr(self.item_end), _abi.ptr(self.item_dst), _abi.ptr(self.item_row),
                                              _abi.ptr(self.long_row), _abi.ptr(self.long_slot0), _abi.ptr(self.long_nslot),
                                              _abi.ptr(self.slot_long), C.c_void_p(self.pass_ptr.ctypes.data),
                                              _abi.stream_ptr()), "plan_build")
        self.n_items, self.n_long, self.n_slots = ni, nl, ns
        self.n_passes = nb
        self.long_cnt = torch.zeros(max(nl, 1), dtype=torch.int32, device=dev)
        self.partial = torch.empty((max(ns, 1), max_d), dtype=torch.float32, device=dev) if ns else None
        self.max_d = max_d

    def _build_stream(self, window):
        """b200rec_stream_build: the pass-major record stream + windows b200rec_spmm_f32_blocked walks"""
        lib = _abi.load()
        dev = self.device
        n_rec, n_win = C.c_int64(0), C.c_int32(0)
        null = C.c_void_p(0)
        args = (_abi.ptr(self.colidx), _abi.ptr(self.vals), self.n_items, _abi.ptr(self.item_start), _abi.ptr(self.item_end),
                _abi.ptr(self.item_dst), C.c_void_p(self.pass_ptr.ctypes.data), self.n_passes, window, C.byref(n_rec),
                C.byref(n_win))
        with torch.cuda.device(dev):
            _abi.check(lib.b200rec_stream_build(*args, null, null, null, _abi.stream_ptr()), "stream_build(size)")
            self.records = torch.empty((max(n_rec.value, 1), 2), dtype=torch.int32, device=dev)
            self.win_start = torch.empty(n_win.value + 1, dtype=torch.int32, device=dev)
            self.pass_win_ptr = np.zeros(self.n_passes + 1, dtype=np.int32)
            _abi.check(lib.b200rec_stream_build(*args, _abi.ptr(self.records), _abi.ptr(self.win_start),
                                                C.c_void_p(self.pass_win_ptr.ctypes.data), _abi.stream_ptr()), "stream_build")
        self.win_counter = torch.zeros(self.n_passes, dtype=torch.int32, device=dev)
        self.n_records, self.n_windows, self.window = n_rec.value, n_win.value, window

    def drop_items(self):
        """the per-segment item arrays of a streamed operand are only needed to build the stream"""
        dev = self.device
        for name in ("item_start", "item_end", "item_dst", "item_row"):
            setattr(self, name, torch.zeros(1, dtype=torch.int32, device=dev))
        self._struct = None

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
            s.n_passes = self.n_passes
            s.pass_ptr = self.pass_ptr.ctypes.data if self.n_passes > 1 else None
            s.records = p(self.records)
            s.win_start = p(self.win_start)
            s.pass_win_ptr = self.pass_win_ptr.ctypes.data if self.pass_win_ptr is not None else None
            s.win_counter = p(self.win_counter)
            self._struct = s
        return self._struct

    # ---- column-blocked execution for tables that do not fit L2 ----
    def blocked_for(self, d):
        """(twin operand with a column-blocked plan, sweep width) for a gathered table of width d, or None"""
        if self.vals is None or self.nbr_scale is not None or self.eid is not None or self.n_passes > 1 \
                or getattr(self, "_row_sliced", False):
            return None
        pol = blocking_policy(self.n_cols, d)
        if pol is None:
            return None
        sweep, block_rows = pol
        key = (sweep, block_rows)
        if key not in self._blocked:
            split = self.phase_split if 0 < self.phase_split < self.n_cols and self.n_rows == self.n_cols else 0
            sides = [(0, self.n_cols)] if not split else [(0, split), (split, self.n_cols)]
            bounds = [0]
            for lo, hi in sides:  # a block never straddles the user | item boundary: each side is gathered by the other
                nblk = max(1, -(-(hi - lo) // block_rows))
                step = -(-(hi - lo) // nblk)
                bounds += [min(hi, lo + (k + 1) * step) for k in range(nblk)]
            bop = CsrOperand(self.rowptr, self.colidx, self.n_cols, vals=self.vals, chunk=self.chunk,
                             max_d=self.max_d, phase_split=self.phase_split, col_bounds=bounds)
            bop.drop_items()
            self._blocked[key] = bop
        return self._blocked[key], sweep

    def hinted_for(self, d):
        """twin of this operand whose gathers carry L2 residency hints, for a gathered table [n_cols, d] larger than L2 (or
        None): the highest-degree source rows of each side -- as many as fit the persisting L2 set-aside -- are flagged
        hot (bit 31 of a private colidx copy) and gathered L2::evict_last; B200REC_HOT_MB sizes the hot set (0 = off),
        B200REC_HOT_COLD = 'first' loads the other rows L2::evict_first.  Same plan, same sums, same bits."""
        import os
        hot_mb = float(os.environ.get("B200REC_HOT_MB", "0"))
        if hot_mb <= 0 or self.vals is None or self.nbr_scale is not None or self.eid is not None or self.n_rows != self.n_cols \
                or getattr(self, "_row_sliced", False):
            return None
        l2 = torch.cuda.get_device_properties(self.device).L2_cache_size
        if self.n_cols * d * 4 <= l2:
            return None
        key = ("hint", d, hot_mb, os.environ.get("B200REC_HOT_COLD", "normal"))
        if key not in self._blocked:
            granted = C.c_int64(0)
            with torch.cuda.device(self.device):
                _abi.check(_abi.load().b200rec_l2_persist(int(hot_mb * (1 << 20)), C.byref(granted)), "l2_persist")
            k = int(min(hot_mb * (1 << 20), granted.value) // (d * 4))
            deg = (self.rowptr[1:] - self.rowptr[:-1]).long()
            hot = torch.zeros(self.n_cols, dtype=torch.bool, device=self.device)
            split = self.phase_split if 0 < self.phase_split < self.n_cols else 0
            for lo, hi in ([(0, self.n_cols)] if not split else [(0, split), (split, self.n_cols)]):
                kk = min(k, hi - lo)
                if kk > 0:
                    hot[lo + torch.topk(deg[lo:hi], kk).indices] = True
            o = object.__new__(CsrOperand)
            o.__dict__.update(self.__dict__)
            o._struct, o._blocked, o._carry = None, {}, {}
            enc = self.colidx.clone()
            enc[hot[self.colidx.long()]] |= -2 ** 31
            o.colidx_enc, o.hot_rows = enc, int(hot.sum())
            o.col_hint = 2 if key[3] == "first" else 1
            o.persist_bytes = granted.value
            self._blocked[key] = o
        return self._blocked[key]

    def carry(self, ld):
        if ld not in self._carry:
            self._carry[ld] = torch.empty((self.n_rows, ld), dtype=torch.float32, device=self.device)
        return self._carry[ld]

    def replan(self, chunk):
        """the same matrix with a different hub-row chunk (own work plan and scratch, shared CSR arrays)"""
        return CsrOperand(self.rowptr, self.colidx, self.n_cols, vals=self.vals, nbr_scale=self.nbr_scale,
                          row_scale=self.row_scale, eid=self.eid, chunk=chunk, max_d=self.max_d, phase_split=self.phase_split)

    def with_scales(self, nbr_scale=None, row_scale=None):
        """same structure/plan, different per-node scale vectors (IGCN anneal: F's values change every epoch)"""
        o = object.__new__(CsrOperand)
        o.__dict__.update(self.__dict__)
        o.nbr_scale, o.row_scale, o._struct = nbr_scale, row_scale, None
        o._blocked, o._carry = {}, {}
        return o

    def row_slice(self, lo, hi=None):
        """rows [lo, hi) -- or a list of such ranges -- as an operand that writes into the FULL output table at global
        row ids (multi-GPU row partition: per-row sums are unchanged, so P-rank results equal 1-rank bit for bit)."""
        ranges = [(lo, hi)] if hi is not None else list(lo)
        o = object.__new__(CsrOperand)
        o.__dict__.update(self.__dict__)
        o._struct = None
        o._blocked, o._carry, o._row_sliced = {}, {}, True
        assert self.n_passes <= 1, "row_slice of a column-blocked plan"
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


def csr_from_coo(rows, cols, n_rows, n_cols, want_first_pos=False):
    """b200rec_csr_build: coalesced int32 CSR of the (row, col) entry list -> (rowptr, colidx, mult or None[, first_pos])"""
    lib = _abi.load()
    rows, cols = rows.contiguous(), cols.contiguous()
    _abi.require_cuda(rows, cols)
    assert rows.dtype == torch.int64 and cols.dtype == torch.int64
    dev = rows.device
    n = rows.numel()
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    colidx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    mult = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if want_first_pos else None
    nnz, dup = C.c_int32(0), C.c_int32(0)
    with torch.cuda.device(dev):
        _abi.check(lib.b200rec_csr_build(_abi.ptr(rows), _abi.ptr(cols), n, 0, 0, n_rows, n_cols, _abi.ptr(rowptr),
                                         _abi.ptr(colidx), _abi.ptr(mult), _abi.ptr(first), C.byref(nnz), C.byref(dup),
                                         _abi.stream_ptr()), "csr_build")
    k = nnz.value
    out = (rowptr, colidx[:k].clone() if k < n else colidx[:k], mult[:k].clone() if dup.value else None)
    return out + (first[:k],) if want_first_pos else out


def build_norm_adj(n_users, n_items, users, items, device=None, chunk=None, d=None):
    """users/items: int64 tensors of the E train pairs (train_array, dataset.py:149-151), any order, duplicates allowed
    (they are summed, like scipy's COO->CSR in utils.py:47-49).  Returns a CsrOperand with .vals / .dinv / .mult.
    Built behind the C ABI: b200rec_adj_build (CSR), b200rec_adj_normalize (values), b200rec_plan_build (work plan)."""
    device = torch.device(device) if device is not None else users.device
    users, items = users.to(device, torch.int64).contiguous(), items.to(device, torch.int64).contiguous()
    n = n_users + n_items
    e = users.numel()
    lib = _abi.load()
    _abi.require_cuda(users, items)
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=device)
    colidx = torch.empty(max(2 * e, 1), dtype=torch.int32, device=device)
    mult = torch.empty(max(2 * e, 1), dtype=torch.float32, device=device)
    nnz, dup = C.c_int32(0), C.c_int32(0)
    with torch.cuda.device(device):
        _abi.check(lib.b200rec_adj_build(_abi.ptr(users), _abi.ptr(items), e, n_users, n_items, _abi.ptr(rowptr),
                                         _abi.ptr(colidx), _abi.ptr(mult), C.byref(nnz), C.byref(dup), _abi.stream_ptr()),
                   "adj_build")
        k = nnz.value
        colidx = colidx[:k].clone() if k < 2 * e else colidx
        mult = mult[:k].clone() if dup.value else None
        dinv = torch.empty(n, dtype=torch.float32, device=device)
        vals = torch.empty(colidx.numel(), dtype=torch.float32, device=device)
        _abi.check(lib.b200rec_adj_normalize(_abi.ptr(rowptr), _abi.ptr(colidx), _abi.ptr(mult), n, _abi.ptr(dinv),
                                             _abi.ptr(vals), _abi.stream_ptr()), "adj_normalize")
    import os
    op = CsrOperand(rowptr, colidx, n, vals=vals, chunk=chunk, d=d,
                    phase_split=0 if os.environ.get("B200REC_NO_PHASE", "0") == "1" else n_users)
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
    rowptr, colidx, mult = csr_from_coo(rows, cols, n, t)
    if mult is None:
        row_sum = (rowptr[1:] - rowptr[:-1]).to(torch.float32)
    else:  # np.sum(feat, axis=1) counts duplicates
        r = torch.repeat_interleave(torch.arange(n, device=device), (rowptr[1:] - rowptr[:-1]).long())
        row_sum = torch.zeros(n, dtype=torch.float32, device=device).index_add_(0, r, mult)
    fwd = CsrOperand(rowptr, colidx, t)
    # transpose: the forward entries re-sorted by (col, row); first_pos = the forward position of every entry
    nnz = colidx.numel()
    r = torch.repeat_interleave(torch.arange(n, device=device), (rowptr[1:] - rowptr[:-1]).long())
    trp, tr, _, eid = csr_from_coo(colidx.long(), r, t, n, want_first_pos=True)
    bwd = CsrOperand(trp, tr, n, eid=eid)
    assert nnz == bwd.nnz
    return FeatOperand(fwd, bwd, row_sum), tu, ti
