"""GPU parity tests, kernel level: libb200rec (through the C ABI) against the CPU oracle and the reference fixtures.
Integer / index outputs must be bit-exact; fp32 results within 1e-5 relative (north_star), and bit-exact against the
C oracle where the summation order is part of the contract (SpMM in CSR order, scores in d order)."""
import numpy as np
import pytest
import torch

from conftest import golden_dataset, golden_model
from oracle import oracle_c as oc
from oracle import ref_port as rp

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", ["lightgcn_tiny", "lightgcn_d128"])
def test_adjacency_structure_bit_exact_values_1ulp(golden, name):
    from b200rec import graph
    g = golden(name)
    users, items = rp.pairs_from_csr(g["train_indptr"], g["train_items"])
    op = graph.build_norm_adj(int(g["n_users"]), int(g["n_items"]), torch.from_numpy(users), torch.from_numpy(items), DEV)
    r, c, v = op.to_coo()
    assert np.array_equal(np.stack([_np(r), _np(c)]), g["adj_idx"])           # CSR construction: bit-exact
    ref = g["adj_val"]
    # numpy's fp32 power is itself 1 ulp off the correctly rounded deg^-1/2 for ~20% of degrees (SURVEY 7 vii), and a
    # value is a product of two of them: allow 3 ulp
    np.testing.assert_allclose(_np(v), ref, rtol=4e-7, atol=0)
    # duplicates are summed like scipy's COO->CSR (utils.py:47-49)
    u2 = np.concatenate([users, users[:50]]); i2 = np.concatenate([items, items[:50]])
    op2 = graph.build_norm_adj(int(g["n_users"]), int(g["n_items"]), torch.from_numpy(u2), torch.from_numpy(i2), DEV)
    a2 = rp.norm_adjacency(int(g["n_users"]), int(g["n_items"]), u2, i2).tocoo()
    r2, c2, v2 = op2.to_coo()
    assert np.array_equal(np.stack([_np(r2), _np(c2)]), np.stack([a2.row, a2.col]))
    np.testing.assert_allclose(_np(v2), a2.data, rtol=3e-7)


@pytest.mark.parametrize("d", [8, 16, 32, 64, 128, 256])
def test_spmm_bit_exact_vs_c_oracle_with_hub_rows(d):
    from b200rec import graph, ops
    rng = np.random.default_rng(d)
    n_rows, n_cols = 700, 6000
    lens = rng.integers(0, 40, n_rows)
    lens[5] = 3000; lens[77] = 1025; lens[300] = 0; lens[699] = 1024   # hubs (split at chunk=1024), empty, exact chunk
    rows = np.repeat(np.arange(n_rows), lens)
    cols = rng.integers(0, n_cols, rows.size)
    key = np.unique(rows.astype(np.int64) * n_cols + cols)
    rows, cols = key // n_cols, key % n_cols
    rp_ = np.zeros(n_rows + 1, dtype=np.int32); np.cumsum(np.bincount(rows, minlength=n_rows), out=rp_[1:])
    vals = rng.standard_normal(rows.size).astype(np.float32)
    x = rng.standard_normal((n_cols, d)).astype(np.float32)
    op = graph.CsrOperand(torch.from_numpy(rp_).to(DEV), torch.from_numpy(cols.astype(np.int32)).to(DEV), n_cols,
                          vals=torch.from_numpy(vals).to(DEV), chunk=1024)
    assert op.n_long >= 1
    xd = torch.from_numpy(x).to(DEV)
    y = torch.empty((n_rows, d), device=DEV)
    add = torch.from_numpy(rng.standard_normal((n_rows, d)).astype(np.float32)).to(DEV)
    out = torch.empty_like(y)
    ops.spmm(op, xd, y=y, addend=add, out=out, out_scale=0.25)
    ref = oc.spmm_csr(rp_, cols, vals, x, chunk=1024)
    assert np.array_equal(_np(y), ref)                                   # same order, same fmaf chain
    assert np.array_equal(_np(out), (_np(add) + ref) * np.float32(0.25))
    # in-place epilogue (out aliases addend), no y
    add2 = add.clone()
    ops.spmm(op, xd, addend=add2, out=add2, out_scale=1.0)
    assert np.array_equal(_np(add2), _np(add) + ref)


@pytest.mark.parametrize("name", ["lightgcn_tiny", "lightgcn_d128"])
def test_propagate_forward_backward(golden, name):
    from b200rec import ops
    g = golden(name)
    ds, m = golden_model(g, name)
    m.eval()
    with torch.no_grad():
        rep = m.get_rep()
    np.testing.assert_allclose(_np(rep), g["rep_eval"], rtol=1e-5, atol=1e-7)
    # backward: dX0 = mean_k A^k G  (A symmetric) -- against autograd through the oracle port
    L = int(g["n_layers"])
    users, items = rp.pairs_from_csr(g["train_indptr"], g["train_items"])
    port = rp.LightGCNPort(int(g["n_users"]), int(g["n_items"]), users, items, g["emb0"], L)
    gout = torch.randn(rep.shape, generator=torch.Generator().manual_seed(1))
    port.get_rep().backward(gout)
    x0 = m.embedding.weight.detach().clone().requires_grad_(True)
    ops.propagate(m.norm_adj, x0, L).backward(gout.to(DEV))
    np.testing.assert_allclose(_np(x0.grad), port.embedding.weight.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_sampler_and_dropout_bit_exact_vs_c_oracle(golden):
    from b200rec import ops
    g = golden("lightgcn_tiny")
    ds = golden_dataset(g)
    ptr, idx = ds.csr("train", device=DEV)
    for step in (0, 1, 12345678901):
        st = torch.tensor([step], dtype=torch.int64, device=DEV)
        got = ops.bpr_sample(ptr, idx, ds.n_users, ds.n_items, 2021, st, 4096)
        ref = oc.bpr_sample(_np(ptr), _np(idx), ds.n_users, ds.n_items, 2021, step, 4096)
        assert np.array_equal(_np(got), ref)
        bits = ops.dropout_bits(100003, 0.3, 77, st)
        assert np.array_equal(_np(bits).view(np.uint32), oc.dropout_mask(100003, 0.3, 77, step))
    # users with empty rows are never drawn
    p2 = torch.tensor([0, 0, 3, 3, 5], dtype=torch.int32, device=DEV)
    i2 = torch.tensor([1, 4, 7, 0, 2], dtype=torch.int32, device=DEV)
    st = torch.zeros(1, dtype=torch.int64, device=DEV)
    b = _np(ops.bpr_sample(p2, i2, 4, 9, 5, st, 2048))
    assert set(np.unique(b[:, 0]).tolist()) == {1, 3}
    assert np.array_equal(b, oc.bpr_sample(_np(p2), _np(i2), 4, 9, 5, 0, 2048))


@pytest.mark.parametrize("d,with_w,reg_mode", [(64, False, 0), (64, False, 1), (128, False, 1), (64, True, 0), (32, True, 0),
                                               (256, False, 1), (16, False, 0), (8, False, 1), (8, True, 0)])
def test_bpr_fused_matches_torch_fp32(d, with_w, reg_mode):
    from b200rec import ops
    gen = torch.Generator(device=DEV).manual_seed(d + reg_mode)
    n_users, n_items, B = 500, 800, 1000  # B not a multiple of the block's samples: ragged tail
    rep = (0.3 * torch.randn((n_users + n_items, d), device=DEV, generator=gen)).requires_grad_(True)
    batch = torch.stack([torch.randint(0, n_users, (B,), device=DEV, generator=gen),
                         torch.randint(0, 40, (B,), device=DEV, generator=gen),      # few items: many duplicate rows
                         torch.randint(0, n_items, (B,), device=DEV, generator=gen)], 1)
    w = (1 + 0.1 * torch.randn(d, device=DEV, generator=gen)).requires_grad_(True) if with_w else None
    l2_reg, scale = 1e-2, 0.37
    u, p, n = rep[batch[:, 0]], rep[n_users + batch[:, 1]], rep[n_users + batch[:, 2]]
    ww = w[None, :] if with_w else 1.0
    loss = torch.nn.functional.softplus((u * n * ww).sum(1) - (u * p * ww).sum(1)).mean()
    if reg_mode == 1:
        loss = loss + l2_reg * ((u * u).sum(1) + (p * p).sum(1) + (n * n).sum(1)).mean()
    (scale * loss).backward()
    g_rep = torch.zeros_like(rep)
    g_w = torch.zeros(d, device=DEV) if with_w else None
    loss_out = torch.zeros(1, device=DEV)
    scratch = ops.bpr_scratch(B, d, DEV)
    for _ in range(2):  # second call: scratch ticket must have reset itself
        g_rep.zero_(); loss_out.zero_()
        ops.bpr_fwd_bwd(rep.detach(), batch, n_users, l2_reg, reg_mode, g_rep, loss_out, scratch,
                        w=w.detach() if with_w else None, g_w=g_w, loss_scale=scale)
    np.testing.assert_allclose(float(loss_out), scale * float(loss), rtol=2e-6)
    np.testing.assert_allclose(_np(g_rep), _np(rep.grad), rtol=1e-4, atol=1e-8)
    if with_w:
        np.testing.assert_allclose(_np(g_w), _np(w.grad), rtol=1e-4, atol=1e-8)


def test_l2_emb0_gather_scatter_adam():
    from b200rec import ops
    gen = torch.Generator(device=DEV).manual_seed(3)
    n_users, n_items, B, d = 300, 400, 777, 64
    emb = (0.2 * torch.randn((n_users + n_items, d), device=DEV, generator=gen)).requires_grad_(True)
    batch = torch.stack([torch.randint(0, n_users, (B,), device=DEV, generator=gen),
                         torch.randint(0, n_items, (B,), device=DEV, generator=gen),
                         torch.randint(0, n_items, (B,), device=DEV, generator=gen)], 1)
    u, p, n = emb[batch[:, 0]], emb[n_users + batch[:, 1]], emb[n_users + batch[:, 2]]
    l2 = (torch.norm(u, p=2, dim=1) ** 2 + torch.norm(p, p=2, dim=1) ** 2 + torch.norm(n, p=2, dim=1) ** 2).mean()
    (1e-3 * l2).backward()
    g = torch.zeros_like(emb); loss = torch.zeros(1, device=DEV)
    ops.bpr_l2_emb0(emb.detach(), batch, n_users, 1e-3, g, loss, ops.bpr_scratch(B, d, DEV))
    np.testing.assert_allclose(float(loss), 1e-3 * float(l2), rtol=2e-6)
    np.testing.assert_allclose(_np(g), _np(emb.grad), rtol=1e-5, atol=1e-9)
    # gather / scatter-add autograd pair == advanced indexing + index_put(accumulate)
    t1 = emb.detach().clone().requires_grad_(True)
    t2 = emb.detach().clone().requires_grad_(True)
    idx = batch[:, 1].contiguous()
    up = torch.randn((B, d), device=DEV, generator=gen)
    ops.gather_rows(t1, idx, n_users).backward(up)
    t2[idx + n_users].backward(up)
    assert torch.equal(ops.gather_rows(t1, idx, n_users), t2[idx + n_users])
    np.testing.assert_allclose(_np(t1.grad), _np(t2.grad), rtol=1e-5, atol=1e-6)
    # Adam == torch.optim.Adam, several steps, n not a multiple of 4
    pw = torch.randn(1003, device=DEV, generator=gen)
    ref = pw.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v, st = torch.zeros_like(pw), torch.zeros_like(pw), torch.zeros(1, dtype=torch.int64, device=DEV)
    for k in range(5):
        gr = torch.randn(1003, device=DEV, generator=gen) * (10.0 ** (k - 2))
        ref.grad = gr.clone(); opt.step()
        ops.adam_step(pw, gr, m, v, st, 1e-3); ops.step_advance(st)
    assert int(st) == 5
    np.testing.assert_allclose(_np(pw), _np(ref), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(_np(m), _np(opt.state[ref]["exp_avg"]), rtol=1e-6, atol=1e-6)  # lerp cancellation
    np.testing.assert_allclose(_np(v), _np(opt.state[ref]["exp_avg_sq"]), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("d,k,n_items", [(64, 20, 500), (64, 100, 3001), (128, 20, 777), (16, 5, 130), (256, 128, 600),
                                         (32, 20, 15)])
def test_score_dense_and_topk_bit_exact_vs_c_oracle(d, k, n_items):
    from b200rec import ops
    rng = np.random.default_rng(d * 1000 + k)
    n_all = 333
    rep_u = (0.1 * rng.standard_normal((n_all, d))).astype(np.float32)
    rep_i = (0.1 * rng.standard_normal((n_items, d))).astype(np.float32)
    rep_i[7] = rep_i[3]; rep_i[min(n_items - 1, 90)] = rep_i[3]          # exact score ties -> id-ascending tie-break
    users = rng.permutation(n_all)[:200].astype(np.int64)
    lens = rng.integers(0, min(30, n_items // 2), n_all)
    lens[users[0]] = max(0, n_items - 3)                                  # fewer than K unmasked items: padded with -1
    ea_ptr = np.zeros(n_all + 1, dtype=np.int32); np.cumsum(lens, out=ea_ptr[1:])
    ea_idx = np.concatenate([np.sort(rng.choice(n_items, l, replace=False)) for l in lens]).astype(np.int32)
    lens_b = rng.integers(0, 5, n_all)
    eb_ptr = np.zeros(n_all + 1, dtype=np.int32); np.cumsum(lens_b, out=eb_ptr[1:])
    eb_idx = np.concatenate([np.sort(rng.choice(n_items, l, replace=False)) for l in lens_b] + [np.zeros(0, int)]).astype(np.int32)
    ru, ri, us = (torch.from_numpy(a).to(DEV) for a in (rep_u, rep_i, users))
    sc = ops.score_dense(ru, us, ri)
    ref_sc = oc.scores(rep_u, users, rep_i)
    assert np.array_equal(_np(sc), ref_sc)
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    for excl_a, excl_b, banned in ((None, None, None), ((ea_ptr, ea_idx), None, None),
                                   ((ea_ptr, ea_idx), (eb_ptr, eb_idx), (n_items // 3, n_items // 2))):
        ids, vals = ops.score_topk(ru, us, ri, k,
                                   excl_a=(t(excl_a[0]), t(excl_a[1])) if excl_a else None,
                                   excl_b=(t(excl_b[0]), t(excl_b[1])) if excl_b else None, banned=banned)
        rid, rval = oc.mask_topk(ref_sc, users, k, excl_a, excl_b, banned or (0, 0))
        assert np.array_equal(_np(ids), rid)
        assert np.array_equal(_np(vals), rval)


def test_hit_matrix():
    from b200rec import ops
    ptr = torch.tensor([0, 2, 2, 5], dtype=torch.int32, device=DEV)
    idx = torch.tensor([3, 9, 0, 4, 8], dtype=torch.int32, device=DEV)
    rec = torch.tensor([[9, 1, 3], [3, 9, 0], [8, -1, 5]], dtype=torch.int32, device=DEV)
    hit = ops.hit_matrix(rec, 0, ptr, idx)
    assert _np(hit).tolist() == [[1, 0, 1], [0, 0, 0], [1, 0, 0]]


def test_spmm_live_list_is_bit_identical(golden):
    """the compacted work list runs exactly the flagged rows' items: same bits on those rows, other rows untouched"""
    from b200rec import graph, ops
    g = golden("lightgcn_tiny")
    users, items = rp.pairs_from_csr(g["train_indptr"], g["train_items"])
    op = graph.build_norm_adj(int(g["n_users"]), int(g["n_items"]), torch.from_numpy(users), torch.from_numpy(items), DEV,
                              chunk=64)   # hub rows split: their pieces must all make the list
    assert op.n_long > 0
    n, d = op.n_rows, 64
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn((n, d), device=DEV, generator=gen)
    acc = torch.randn((n, d), device=DEV, generator=gen)
    flags = (torch.rand(n, device=DEV, generator=gen) < 0.2).to(torch.uint8)
    flags[op.long_row[:op.n_long].long()] = 1
    full = torch.empty_like(x)
    ops.spmm(op, x, addend=acc, out=full, out_scale=0.25)
    live = torch.empty(op.n_items, dtype=torch.int32, device=DEV)
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.live_items(op, flags, live, cnt)
    n_live = int(cnt)
    want = flags[op.item_row[:op.n_items].long()].bool()
    assert n_live == int(want.sum())
    assert sorted(_np(live[:n_live]).tolist()) == torch.nonzero(want).flatten().tolist()
    out = torch.full_like(x, 7.0)
    ops.spmm_live(op, x, live, cnt, n_live + 5, addend=acc, out=out, out_scale=0.25)
    sel = flags.bool()
    assert torch.equal(out[sel], full[sel]) and bool((out[~sel] == 7.0).all())
    cnt.zero_()                                         # empty list: nothing is written
    out.fill_(7.0)
    ops.spmm_live(op, x, live, cnt, 16, addend=acc, out=out, out_scale=0.25)
    assert bool((out == 7.0).all())


@pytest.mark.parametrize("n_users,k,topks", [(1000, 20, [1, 5, 10, 15, 20]), (333, 100, list(range(5, 101, 5)) + [1]), (5, 3, [3])])
def test_rank_metrics_kernel_matches_oracle(n_users, k, topks):
    """fused Precision / Recall / NDCG pass vs the restated calculate_metrics (oracle/ref_port.py, trainer.py:115-144)"""
    from b200rec import ops
    rng = np.random.default_rng(n_users + k)
    n_items = 400
    lens = rng.integers(0, 30, n_users)
    lens[:3] = [0, 1, 29][:min(3, n_users)]          # users without eval items are excluded from the means
    eval_data = [sorted(rng.choice(n_items, l, replace=False).tolist()) for l in lens]
    rec = np.stack([rng.permutation(n_items)[:k] for _ in range(n_users)]).astype(np.int32)
    rec[n_users // 2, k - 1] = -1                    # an unfilled slot never hits
    ptr = np.zeros(n_users + 1, dtype=np.int32)
    np.cumsum(lens, out=ptr[1:])
    idx = np.concatenate([np.asarray(x, dtype=np.int32) for x in eval_data] + [np.zeros(0, np.int32)])
    ks = sorted(set(topks))
    sums = _np(ops.rank_metrics(torch.from_numpy(rec).to(DEV), 0, torch.from_numpy(ptr).to(DEV),
                                torch.from_numpy(idx if idx.size else np.zeros(1, np.int32)).to(DEV), ks))
    ref = rp.calculate_metrics(eval_data, rec, ks)
    assert sums[-1] == (lens > 0).sum()
    for i, kk in enumerate(ks):
        for j, name in enumerate(("Precision", "Recall", "NDCG")):
            assert abs(sums[3 * i + j] / sums[-1] - float(ref[name][kk])) < 2e-6, (name, kk)


def test_spmm_row_and_source_masks_are_bit_identical(golden):
    """dst_flags / src_flags skip work without changing a bit of the rows that are computed"""
    from b200rec import graph, ops
    g = golden("lightgcn_tiny")
    users, items = rp.pairs_from_csr(g["train_indptr"], g["train_items"])
    op = graph.build_norm_adj(int(g["n_users"]), int(g["n_items"]), torch.from_numpy(users), torch.from_numpy(items), DEV,
                              chunk=64)   # small chunk: hub rows are split, so the piece -> row lookup is exercised
    assert op.n_long > 0
    n, d = op.n_rows, 64
    gen = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn((n, d), device=DEV, generator=gen)
    full = torch.empty_like(x)
    ops.spmm(op, x, y=full)
    flags = (torch.rand(n, device=DEV, generator=gen) < 0.2).to(torch.uint8)
    part = torch.full_like(x, 7.0)
    ops.spmm(op, x, y=part, dst_flags=flags)
    sel = flags.bool()
    assert torch.equal(part[sel], full[sel]) and bool((part[~sel] == 7.0).all())
    xs = x * sel[:, None]                      # source rows outside the mask are zero
    ref = torch.empty_like(x); ops.spmm(op, xs, y=ref)
    got = torch.empty_like(x); ops.spmm(op, xs, y=got, src_flags=flags)
    assert torch.equal(got, ref)
    # the propagation entry points: restricted last layer / sparse first hop
    L = 3
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    m_full, m_part = torch.empty_like(x), torch.empty_like(x)
    ops.propagate_fwd(op, x, L, bufs, m_full)
    ops.propagate_fwd(op, x, L, bufs, m_part, needed_rows=flags)
    assert torch.equal(m_part[sel], m_full[sel])
    d_full, d_part = torch.empty_like(x), torch.empty_like(x)
    ops.propagate_bwd(op, xs, L, bufs, d_full)
    ops.propagate_bwd(op, xs, L, bufs, d_part, nonzero_rows=flags)
    assert torch.equal(d_part, d_full)


@pytest.mark.parametrize("d", [16, 64, 128])
def test_bpr_ordered_scatter_is_deterministic_and_matches_atomic(d):
    """heavy duplicate ids (popular items): the ordered scatter forms each touched row once, in slot order -> the same bits
    on every run, and the same values as the per-sample atomic scatter / a float64 restatement"""
    from b200rec import ops
    rng = np.random.default_rng(d)
    nu, ni, b = 50, 40, 2048
    rep = torch.from_numpy(rng.standard_normal((nu + ni, d)).astype(np.float32)).to(DEV)
    batch_np = np.stack([rng.integers(0, nu, b), rng.integers(0, 7, b), rng.integers(0, ni, b)], 1).astype(np.int64)
    batch = torch.from_numpy(batch_np).to(DEV)
    scratch = ops.bpr_scratch(b, d, DEV)
    loss_a = torch.zeros(1, device=DEV)
    g_atomic = torch.zeros_like(rep)
    ops.bpr_fwd_bwd(rep, batch, nu, 1e-3, 1, g_atomic, loss_a, scratch)
    grp = ops.bpr_grouping(b, DEV)
    ops.bpr_group_rows(batch, nu, grp)
    rows = np.concatenate([batch_np[:, :1], batch_np[:, 1:] + nu], 1).reshape(-1)
    # order = the slots in (row, slot) order (exactly numpy's stable argsort of the rows); bit 31 marks a continued row
    raw = grp[0].cpu().numpy().view(np.uint32)
    order, cont = (raw & 0x7fffffff).astype(np.int64), (raw >> 31).astype(bool)
    expect = np.argsort(rows, kind="stable")
    assert np.array_equal(order, expect)
    assert np.array_equal(cont, np.concatenate([[False], rows[expect][1:] == rows[expect][:-1]]))
    assert int((~cont).sum()) == np.unique(rows).size
    outs = []
    for _ in range(3):
        g = torch.zeros_like(rep)
        loss_o = torch.zeros(1, device=DEV)
        ops.bpr_fwd_bwd_ordered(rep, batch, nu, 1e-3, 1, g, loss_o, scratch, grp)
        outs.append(g)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert float(loss_o) == float(loss_a)
    torch.testing.assert_close(outs[0], g_atomic, rtol=2e-5, atol=1e-6)
    # float64 restatement (trainer.py:414-424 + its autograd)
    r64 = rep.double().cpu().numpy()
    u, p, n = r64[batch_np[:, 0]], r64[nu + batch_np[:, 1]], r64[nu + batch_np[:, 2]]
    x = (u * n).sum(1) - (u * p).sum(1)
    coef = (1 / (1 + np.exp(-x)) / b)[:, None]
    rc = 2 * 1e-3 / b
    ref = np.zeros_like(r64)
    np.add.at(ref, batch_np[:, 0], coef * (n - p) + rc * u)
    np.add.at(ref, nu + batch_np[:, 1], -coef * u + rc * p)
    np.add.at(ref, nu + batch_np[:, 2], coef * u + rc * n)
    np.testing.assert_allclose(outs[0].cpu().numpy(), ref, rtol=1e-4, atol=1e-6)
    # accumulate mode + weighted scores (IGCN auxiliary loss): added on top of what the table holds
    w = torch.from_numpy(rng.standard_normal(d).astype(np.float32)).to(DEV)
    base = torch.from_numpy(rng.standard_normal((nu + ni, d)).astype(np.float32)).to(DEV)
    ga, gw_a = base.clone(), torch.zeros(d, device=DEV)
    ops.bpr_fwd_bwd(rep, batch, nu, 0.0, 0, ga, loss_a, scratch, w=w, g_w=gw_a, loss_scale=0.3)
    go, gw_o = base.clone(), torch.zeros(d, device=DEV)
    ops.bpr_fwd_bwd_ordered(rep, batch, nu, 0.0, 0, go, loss_o, scratch, grp, w=w, g_w=gw_o, loss_scale=0.3, accumulate=True)
    torch.testing.assert_close(go, ga, rtol=2e-5, atol=2e-6)
    assert torch.equal(gw_a, gw_o)
    # layer-0 regulariser and the row clear
    e1, e2 = base.clone(), base.clone()
    ops.bpr_l2_emb0(rep, batch, nu, 1e-2, e1, loss_a, scratch)
    ops.bpr_l2_emb0_ordered(rep, batch, nu, 1e-2, e2, loss_o, scratch, grp)
    torch.testing.assert_close(e1, e2, rtol=2e-5, atol=2e-6)
    # folded form the LightGCN step uses: the loss term rides the BPR launch, gradient + row clear are one launch
    loss_sep, loss_fold = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    g_sep, g_fold = torch.zeros_like(rep), torch.zeros_like(rep)
    ops.bpr_fwd_bwd_ordered(rep, batch, nu, 0.0, 0, g_sep, loss_sep, scratch, grp)
    e3 = base.clone()
    ops.bpr_l2_emb0_ordered(base, batch, nu, 1e-2, e3, loss_sep, scratch, grp)
    ops.bpr_fwd_bwd_ordered(rep, batch, nu, 0.0, 0, g_fold, loss_fold, scratch, grp, emb0=base, l2_emb0=1e-2)
    e4 = base.clone()
    ops.bpr_l2_emb0_ordered(base, batch, nu, 1e-2, e4, None, None, grp, clear_table=g_fold)
    assert torch.equal(e3, e4)
    assert not bool(g_fold.any())                                   # every row the batch touched is zero again
    torch.testing.assert_close(loss_fold, loss_sep, rtol=1e-5, atol=1e-7)
    ops.clear_rows(batch, nu, e2)
    touched = torch.zeros(nu + ni, dtype=torch.bool, device=DEV)
    touched[torch.from_numpy(rows).to(DEV)] = True
    assert bool((e2[touched] == 0).all()) and torch.equal(e2[~touched], base[~touched])
    from b200rec import _abi
    with pytest.raises(_abi.B200RecError):
        ops.bpr_group_rows(torch.zeros((4097, 3), dtype=torch.int64, device=DEV), nu, ops.bpr_grouping(4097, DEV))
    # a batch of the largest supported size, with a ragged tail (3B not a multiple of the warp's slots), still ranks exactly
    for bb in (4096, 1, 333):
        bt = torch.from_numpy(np.stack([rng.integers(0, nu, bb), rng.integers(0, 7, bb), rng.integers(0, ni, bb)], 1)).to(DEV)
        g2 = ops.bpr_grouping(bb, DEV)
        ops.bpr_group_rows(bt, nu, g2)
        r2 = (bt.cpu().numpy() + np.array([0, nu, nu])).reshape(-1)
        assert np.array_equal(g2[0].cpu().numpy().view(np.uint32) & 0x7fffffff, np.argsort(r2, kind="stable"))
