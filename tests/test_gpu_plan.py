"""GPU: the device-side operand builders (b200rec_csr_build / b200rec_adj_build / b200rec_plan_build) and the
column-blocked SpMM (b200rec_spmm_f32_blocked) against scipy, the plan's stated invariants and the C oracle."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from b200rec import graph, ops
from oracle import oracle_c as oc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _powerlaw_lens(rng, n, hubs):
    lens = rng.integers(0, 40, size=n)
    lens[rng.choice(n, size=hubs, replace=False)] = rng.integers(300, 3000, size=hubs)
    lens[rng.choice(n, size=max(1, n // 10), replace=False)] = 0
    return lens


def _random_csr(rng, n_rows, n_cols, hubs=3):
    lens = np.minimum(_powerlaw_lens(rng, n_rows, hubs), n_cols)
    rp = np.zeros(n_rows + 1, dtype=np.int32)
    np.cumsum(lens, out=rp[1:])
    cols = np.concatenate([np.sort(rng.choice(n_cols, size=int(k), replace=False)) for k in lens] + [np.zeros(0, np.int64)])
    return rp, cols.astype(np.int32)


@pytest.mark.parametrize("seed", [0, 1])
def test_csr_build_matches_scipy_with_duplicates(seed):
    rng = np.random.default_rng(seed)
    n_rows, n_cols, n = 300, 170, 5000
    rows, cols = rng.integers(0, n_rows, n), rng.integers(0, n_cols, n)
    ref = sp.coo_matrix((np.ones(n, np.float32), (rows, cols)), shape=(n_rows, n_cols)).tocsr()
    ref.sort_indices()
    rp, ci, mult, first = graph.csr_from_coo(torch.from_numpy(rows).to(DEV), torch.from_numpy(cols).to(DEV), n_rows, n_cols,
                                             want_first_pos=True)
    assert np.array_equal(rp.cpu().numpy(), ref.indptr) and np.array_equal(ci.cpu().numpy(), ref.indices)
    assert mult is not None and np.array_equal(mult.cpu().numpy(), ref.data)
    fp = first.cpu().numpy()
    r_of = np.repeat(np.arange(n_rows), np.diff(ref.indptr))
    assert np.array_equal(rows[fp], r_of) and np.array_equal(cols[fp], ref.indices)
    # first_pos is the FIRST list position of each merged pair
    key = rows * n_cols + cols
    first_of = {}
    for i, k in enumerate(key.tolist()):
        first_of.setdefault(k, i)
    assert fp.tolist() == [first_of[int(r) * n_cols + int(c)] for r, c in zip(r_of, ref.indices)]
    # no duplicates -> mult is None
    uniq = np.unique(key)
    rp2, ci2, mult2 = graph.csr_from_coo(torch.from_numpy(uniq // n_cols).to(DEV), torch.from_numpy(uniq % n_cols).to(DEV),
                                         n_rows, n_cols)
    assert mult2 is None and ci2.numel() == uniq.size and np.array_equal(rp2.cpu().numpy(), ref.indptr)


def test_csr_build_rejects_out_of_range():
    from b200rec import _abi
    with pytest.raises(_abi.B200RecError):
        graph.csr_from_coo(torch.tensor([0, 5], device=DEV), torch.tensor([0, 1], device=DEV), 5, 4)


def _check_plan(op, rp, cols, chunk, bounds, row_order=False):
    n_rows = len(rp) - 1
    ni = op.n_items
    start, end, dst, row = (t[:ni].cpu().numpy() for t in (op.item_start, op.item_end, op.item_dst, op.item_row))
    lens = np.diff(rp)
    cover = np.zeros(int(rp[-1]), dtype=np.int32)
    for s, e, r in zip(start, end, row):
        assert rp[r] <= s <= e <= rp[r + 1] and e - s <= chunk
        cover[s:e] += 1
    assert (cover == 1).all()
    hub = dst < 0
    code = np.where(hub, ~dst, dst).astype(np.int64)
    ident, not_first, not_last = code & 0x1fffffff, (code >> 29) & 1, (code >> 30) & 1
    long_rows = np.nonzero(lens > chunk)[0]
    assert op.n_long == len(long_rows) and sorted(op.long_row[:op.n_long].cpu().tolist()) == long_rows.tolist()
    assert op.n_slots == int(sum(-(-int(lens[r]) // chunk) for r in long_rows))
    assert np.array_equal(ident[~hub], row[~hub]) and not np.isin(row[~hub], long_rows).any()
    assert sorted(set(row[~hub].tolist())) == np.nonzero(lens <= chunk)[0].tolist()   # empty rows too
    lrow, lslot0, lnslot = (t[:op.n_long].cpu().numpy() for t in (op.long_row, op.long_slot0, op.long_nslot))
    slot_long = op.slot_long[:op.n_slots].cpu().numpy()
    for li, r in enumerate(lrow):
        assert lnslot[li] == -(-int(lens[r]) // chunk) and (slot_long[lslot0[li]:lslot0[li] + lnslot[li]] == li).all()
        for q in range(lnslot[li]):  # piece q covers exactly [rp[r] + q*chunk, ...)
            idx = np.nonzero(hub & (ident == lslot0[li] + q))[0]
            assert start[idx].min() == rp[r] + q * chunk and end[idx].max() == min(rp[r] + (q + 1) * chunk, rp[r + 1])
    # every virtual row (plain row / hub piece): consecutive slices, flags mark continuation
    vrow = np.where(hub, -1 - ident, ident)
    for v in np.unique(vrow):
        idx = np.nonzero(vrow == v)[0]
        idx = idx[np.argsort(start[idx], kind="stable")]
        assert (start[idx][1:] == end[idx][:-1]).all()
        assert not_first[idx].tolist() == [0] + [1] * (len(idx) - 1)
        assert not_last[idx].tolist() == [1] * (len(idx) - 1) + [0]
    # passes
    pp = op.pass_ptr
    assert pp[0] == 0 and pp[-1] == ni and (np.diff(pp) >= 0).all()
    if bounds is None:
        assert op.n_passes == 1 and not not_first.any() and not not_last.any()
    else:
        for b in range(len(bounds) - 1):
            for k in range(pp[b], pp[b + 1]):
                c = cols[start[k]:end[k]]
                assert c.size == 0 or (c.min() >= bounds[b] and c.max() < bounds[b + 1])
            if row_order:
                assert (np.diff(start[pp[b]:pp[b + 1]]) >= 0).all()         # CSR order = rows ascending inside a pass
            else:
                assert (np.diff((end - start)[pp[b]:pp[b + 1]]) <= 0).all()   # longest first inside a pass


@pytest.mark.parametrize("chunk", [32, 256])
@pytest.mark.parametrize("nblocks", [1, 2, 7])
def test_plan_build_invariants(chunk, nblocks):
    rng = np.random.default_rng(chunk + nblocks)
    n_rows, n_cols = 400, 3500
    rp, cols = _random_csr(rng, n_rows, n_cols)
    bounds = None if nblocks == 1 else np.linspace(0, n_cols, nblocks + 1).astype(np.int32)
    op = graph.CsrOperand(torch.from_numpy(rp).to(DEV), torch.from_numpy(cols).to(DEV), n_cols, chunk=chunk, col_bounds=bounds,
                          row_order=0)
    _check_plan(op, rp, cols, chunk, bounds)
    if bounds is not None:  # row order (the record stream's input): same items, rows ascending inside a pass
        op1 = graph.CsrOperand(torch.from_numpy(rp).to(DEV), torch.from_numpy(cols).to(DEV), n_cols, chunk=chunk,
                               col_bounds=bounds, row_order=1)
        _check_plan(op1, rp, cols, chunk, bounds, row_order=True)
    if nblocks == 1:  # two scheduling phases of a single-pass plan
        op2 = graph.CsrOperand(torch.from_numpy(rp).to(DEV), torch.from_numpy(cols).to(DEV), n_cols, chunk=chunk, phase_split=150)
        row = op2.item_row[:op2.n_items].cpu().numpy()
        ln = (op2.item_end - op2.item_start)[:op2.n_items].cpu().numpy()
        first = row < 150
        assert first[:first.sum()].all() and not first[first.sum():].any()
        for ph in (first, ~first):
            assert (np.diff(ln[ph]) <= 0).all()
        _check_plan(op2, rp, cols, chunk, None)


def test_record_stream_structure():
    """the pass-major stream: per item [carry-in?] edges [end], windows hold whole items and cover every record once"""
    rng = np.random.default_rng(3)
    n = 700
    rp, cols = _random_csr(rng, n, n, hubs=4)
    vals = rng.standard_normal(cols.size).astype(np.float32)
    bounds = np.array([0, 90, 300, 301, 650, n], dtype=np.int32)
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    op = graph.CsrOperand(t(rp), t(cols), n, vals=t(vals), chunk=64, col_bounds=bounds, window=32)
    rec = op.records.cpu().numpy()
    x, w = rec[:, 0].view(np.uint32), rec[:, 1].view(np.float32)
    typ, hub, ident = x >> 30, (x >> 29) & 1, x & 0x1fffffff
    ni = op.n_items
    start, end, dst = (a[:ni].cpu().numpy() for a in (op.item_start, op.item_end, op.item_dst))
    code = np.where(dst < 0, ~dst, dst).astype(np.int64)
    nf, nl = (code >> 29) & 1, (code >> 30) & 1
    assert op.n_records == int((end - start).sum() + nf.sum() + ni) == rec.shape[0]
    o = 0
    for i in range(ni):
        if nf[i]:
            assert typ[o] == 1 and w[o] == 1.0 and ident[o] == (code[i] & 0x1fffffff) and hub[o] == (dst[i] < 0)
            o += 1
        k = end[i] - start[i]
        assert (typ[o:o + k] == 0).all() and np.array_equal(ident[o:o + k], cols[start[i]:end[i]])
        assert np.array_equal(w[o:o + k], vals[start[i]:end[i]])
        o += k
        assert typ[o] == (2 if nl[i] else 3) and ident[o] == (code[i] & 0x1fffffff) and hub[o] == (dst[i] < 0)
        o += 1
    assert o == rec.shape[0]
    ws, pw = op.win_start.cpu().numpy(), op.pass_win_ptr
    assert ws[0] == 0 and ws[-1] == rec.shape[0] and (np.diff(ws) >= 0).all() and pw[-1] == len(ws) - 1
    item_first = np.concatenate([[0], np.cumsum((end - start) + nf + 1)])
    assert np.isin(ws, item_first).all()                      # windows start at item boundaries
    for b in range(op.n_passes):                               # and never straddle a pass
        assert ws[pw[b]] == item_first[op.pass_ptr[b]] and ws[pw[b + 1]] == item_first[op.pass_ptr[b + 1]]


@pytest.mark.parametrize("d", [16, 64, 128])
def test_column_ordered_hub_chunks_same_plan_same_bits(d, monkeypatch):
    """B200REC_PLAN_COLSORT=1 (the default for graphs of >= 16 M entries): the full chunks of hub rows are scheduled by their
    first source column instead of row by row.  Same work items (the plan's invariants hold), full chunks in column order,
    and the product equals the row-ordered plan's and the C oracle's bit for bit."""
    rng = np.random.default_rng(100 + d)
    n, chunk = 1200, 64
    rp, cols = _random_csr(rng, n, n, hubs=12)
    vals = rng.standard_normal(cols.size).astype(np.float32)
    x = rng.standard_normal((n, d)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    ops_ = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("B200REC_PLAN_COLSORT", flag)
        ops_[flag] = graph.CsrOperand(t(rp), t(cols), n, vals=t(vals), chunk=chunk, phase_split=500)
        _check_plan(ops_[flag], rp, cols, chunk, None)
    a, b = ops_["0"], ops_["1"]
    key = lambda o: sorted(zip(*(v[:o.n_items].cpu().tolist() for v in (o.item_start, o.item_end, o.item_dst, o.item_row))))  # noqa: E731
    assert key(a) == key(b)                                       # the same work items, in another order
    start, end, row = (v[:b.n_items].cpu().numpy() for v in (b.item_start, b.item_end, b.item_row))
    assert not np.array_equal(start, a.item_start[:a.n_items].cpu().numpy())
    for ph in (row < 500, row >= 500):                            # inside a phase: full chunks first, by first column
        full = ph & (end - start == chunk) & np.isin(row, np.nonzero(np.diff(rp) > chunk)[0])
        first_col = cols[start[full]]
        assert full.sum() > 10 and (np.diff(first_col) >= 0).all()
        assert (np.diff((end - start)[ph]) <= 0).all()
    xd = t(x)
    ya, yb = torch.empty_like(xd), torch.empty_like(xd)
    for _ in range(2):                                            # twice: the hub counters are reused
        ops.spmm(a, xd, y=ya)
        ops.spmm(b, xd, y=yb)
    torch.cuda.synchronize()
    assert torch.equal(ya, yb)
    assert np.array_equal(yb.cpu().numpy(), oc.spmm_csr(rp, cols, vals, x, chunk=chunk))


@pytest.mark.parametrize("d,sweep", [(16, 16), (64, 64), (128, 64), (128, 32), (256, 256)])
def test_blocked_spmm_bit_identical_to_single_pass_and_oracle(d, sweep):
    rng = np.random.default_rng(d + sweep)
    n = 900
    rp, cols = _random_csr(rng, n, n, hubs=5)
    vals = rng.standard_normal(cols.size).astype(np.float32)
    x = rng.standard_normal((n, d)).astype(np.float32)
    add = rng.standard_normal((n, d)).astype(np.float32)
    chunk = 64
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    op = graph.CsrOperand(t(rp), t(cols), n, vals=t(vals), chunk=chunk)
    bounds = np.array([0, 100, 130, 131, 500, 500, 777, n], dtype=np.int32)   # uneven, one empty, one single-row block
    bop = graph.CsrOperand(t(rp), t(cols), n, vals=t(vals), chunk=chunk, col_bounds=bounds, window=32 + 32 * (d % 3))
    xd, addd = t(x), t(add)
    y1, o1 = torch.empty_like(xd), torch.empty_like(xd)
    ops.spmm(op, xd, y=y1, addend=addd, out=o1, out_scale=0.2)
    y2, o2 = torch.full_like(xd, 7.0), torch.full_like(xd, 7.0)
    carry = torch.full((n, d), float("nan"), device=DEV)
    ops.spmm_blocked(bop, xd, sweep, carry, y=y2, addend=addd, out=o2, out_scale=0.2)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2) and torch.equal(o1, o2)
    ref = oc.spmm_csr(rp, cols, vals, x, chunk=chunk)
    assert np.array_equal(y2.cpu().numpy(), ref)
    # in place running sum (out aliases addend), repeated launches reuse the hub counters
    acc1, acc2 = addd.clone(), addd.clone()
    for _ in range(2):
        ops.spmm(op, xd, addend=acc1, out=acc1)
        ops.spmm_blocked(bop, xd, sweep, carry, addend=acc2, out=acc2)
    assert torch.equal(acc1, acc2)


def test_blocking_policy_and_dispatch(monkeypatch):
    """ops.spmm switches to the blocked plan when the table outgrows the block budget, with identical bits"""
    rng = np.random.default_rng(5)
    nu, ni, e = 700, 500, 9000
    users, items = rng.integers(0, nu, e), rng.integers(0, ni, e)
    monkeypatch.setenv("B200REC_BLOCK_MB", "0")
    a = graph.build_norm_adj(nu, ni, torch.from_numpy(users), torch.from_numpy(items), device=DEV, chunk=64)
    x = torch.randn(nu + ni, 128, device=DEV)
    y0 = torch.empty_like(x)
    ops.spmm(a, x, y=y0)
    assert a.blocked_for(128) is None
    monkeypatch.setenv("B200REC_BLOCK_MAX_D", "128")
    monkeypatch.setenv("B200REC_SWEEP_D", "64")
    monkeypatch.setenv("B200REC_BLOCK_MB", "0.05")   # 51 KiB blocks, 64-column sweeps
    blk = a.blocked_for(128)
    assert blk is not None and blk[1] == 64 and blk[0].n_passes >= 2
    b = np.asarray(blk[0].col_bounds)
    assert nu in b.tolist()                           # no block straddles the user | item boundary
    y1 = torch.empty_like(x)
    ops.spmm(a, x, y=y1)
    assert torch.equal(y0, y1)
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    m0, m1 = torch.empty_like(x), torch.empty_like(x)
    ops.propagate_fwd(a, x, 3, bufs, m1)
    monkeypatch.setenv("B200REC_BLOCK_MB", "0")
    ops.propagate_fwd(a, x, 3, bufs, m0)
    assert torch.equal(m0, m1)


def test_l2_residency_hints_are_bit_identical(monkeypatch):
    """a table larger than L2 with the top-degree source rows flagged hot (bit 31 of a private colidx copy, evict_last into
    the persisting set-aside): same plan, same sums, same bits as the plain operand"""
    import ctypes as C
    from b200rec import _abi
    rng = np.random.default_rng(9)
    nu, ni, e = 200_000, 140_000, 2_000_000
    users = torch.from_numpy(rng.integers(0, nu, e))
    items = torch.from_numpy((rng.pareto(1.2, e) * 50).astype(np.int64) % ni)     # skewed item popularity
    a = graph.build_norm_adj(nu, ni, users, items, device=DEV)
    x = torch.randn(nu + ni, 128, device=DEV)
    assert x.numel() * 4 > torch.cuda.get_device_properties(0).L2_cache_size
    monkeypatch.setenv("B200REC_HOT_MB", "0")
    assert a.hinted_for(128) is None
    y0, acc0 = torch.empty_like(x), torch.ones_like(x)
    ops.spmm(a, x, y=y0, addend=acc0, out=acc0)
    try:
        for cold in ("normal", "first"):
            monkeypatch.setenv("B200REC_HOT_MB", "24")
            monkeypatch.setenv("B200REC_HOT_COLD", cold)
            h = a.hinted_for(128)
            assert h is not None and h.col_hint == (2 if cold == "first" else 1) and 0 < h.hot_rows <= 2 * (24 << 20) // 512
            assert h.persist_bytes >= 16 << 20 and bool((h.colidx_enc < 0).any())
            y1, acc1 = torch.empty_like(x), torch.ones_like(x)
            ops.spmm(a, x, y=y1, addend=acc1, out=acc1)
            assert torch.equal(y0, y1) and torch.equal(acc0, acc1)
            assert a.hinted_for(64) is None          # 87 MB table: fits L2, left alone
    finally:
        _abi.check(_abi.load().b200rec_l2_persist(0, C.c_void_p(0)), "l2_persist(0)")
