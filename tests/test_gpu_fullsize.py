"""GPU: size-independent properties at the BASELINE.json sizes, where the CPU oracle would take minutes to hours.
C2 = Yelp2018-shaped (31 668 x 38 048, 1.56 M train edges, D=64, L=3); C4 = 2M x 1M, 100 M edges, D=128, L=4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def c2():
    import dataset as D
    import model as M
    from b200rec import synth
    g = synth.generate_named("c2", seed=0, device=DEV)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": g})
    torch.manual_seed(0)
    m = M.get_model({"name": "LightGCN", "embedding_size": 64, "n_layers": 3, "device": DEV}, ds)
    return ds, m


def _dot(a, b):
    return float((a.double() * b.double()).sum())


def test_c2_graph_invariants(c2):
    ds, m = c2
    a = m.norm_adj
    assert a.nnz == 2 * len(ds) and a.n_rows == ds.n_users + ds.n_items
    r, c, v = a.to_coo()
    assert bool((r[1:] * a.n_cols + c[1:] > r[:-1] * a.n_cols + c[:-1]).all())      # row-major sorted, no duplicates
    assert bool(((r < ds.n_users) ^ (c < ds.n_users)).all())                          # bipartite: users <-> items only
    # bit-wise symmetric values: the transpose has the same value multiset at mirrored positions
    key = c * a.n_cols + r
    order = torch.argsort(key)
    assert torch.equal(v[order], v)
    deg = (a.rowptr[1:] - a.rowptr[:-1]).float().clamp(min=1)
    np.testing.assert_allclose(a.dinv.cpu().numpy(), (deg ** -0.5).cpu().numpy(), rtol=2e-7)


def test_c2_propagation_linearity_and_adjoint(c2):
    from b200rec import ops
    ds, m = c2
    a, n, L = m.norm_adj, m.n_users + m.n_items, 3
    gen = torch.Generator(device=DEV).manual_seed(1)
    x, y = (torch.randn((n, 64), device=DEV, generator=gen) for _ in range(2))
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    px, py, pxy = (torch.empty_like(x) for _ in range(3))
    ops.propagate_fwd(a, x, L, bufs, px)
    ops.propagate_fwd(a, y, L, bufs, py)
    ops.propagate_fwd(a, 2.0 * x - 0.5 * y, L, bufs, pxy)
    torch.testing.assert_close(pxy, 2.0 * px - 0.5 * py, rtol=1e-4, atol=1e-5)       # linearity
    by = torch.empty_like(x)
    ops.propagate_bwd(a, y, L, bufs, by)
    assert abs(_dot(px, y) - _dot(x, by)) <= 1e-5 * abs(_dot(px, y)) + 1e-6          # <P x, y> == <x, P^T y>
    torch.testing.assert_close(by, py, rtol=1e-4, atol=1e-5)                          # P is symmetric (A is)
    # determinism: same bits on a second launch (ordered hub reduction, no atomics)
    px2 = torch.empty_like(x)
    ops.propagate_fwd(a, x, L, bufs, px2)
    assert torch.equal(px, px2)


def test_c2_sampler_and_training_epoch(c2):
    import trainer as T
    ds, m = c2
    tr = T.get_trainer({"name": "BPRTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4, "device": DEV, "n_epochs": 1,
                        "batch_size": 2048, "dataloader_num_workers": 0, "test_batch_size": 512,
                        "topks": [1, 5, 10, 15, 20]}, ds, m)
    m.train()
    eng = tr._engine()
    ptr, idx = ds.csr("train", device=DEV)
    for _ in range(3):
        eng.step()
        b = eng.batch
        u, p, ng = b[:, 0], b[:, 1], b[:, 2]
        assert bool(((ptr[u + 1] - ptr[u]) > 0).all())
        key = ptr.long()  # membership by searching the user's sorted row
        lo = key[u]
        flat = idx.long()
        pos_hit = torch.zeros_like(u, dtype=torch.bool)
        neg_hit = torch.zeros_like(u, dtype=torch.bool)
        for k in range(int((ptr[1:] - ptr[:-1]).max())):
            inside = lo + k < key[u + 1]
            cur = flat[(lo + k).clamp(max=flat.numel() - 1)]
            pos_hit |= inside & (cur == p)
            neg_hit |= inside & (cur == ng)
            if k > 4000:
                break
        assert bool(pos_hit.all()) and not bool(neg_hit.any())
    l0 = tr.train_one_epoch()
    l1 = tr.train_one_epoch()
    assert np.isfinite(l0) and l1 < l0 < 0.6932                                       # BPR loss starts at ln 2 and falls


def test_c2_topk_properties(c2):
    import trainer as T
    ds, m = c2
    tr = T.get_trainer({"name": "BPRTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4, "device": DEV, "n_epochs": 1,
                        "batch_size": 2048, "dataloader_num_workers": 0, "test_batch_size": 512,
                        "topks": [1, 5, 10, 15, 20]}, ds, m)
    recs = {}
    for prec in (0, 1):
        tr.eval_precision = prec
        recs[prec] = tr.recommend_all("test")
    assert torch.equal(recs[0], recs[1])                                              # tcgen05 path == exact path, 31 668 users
    rec = recs[0].long()
    assert bool((rec >= 0).all()) and bool((rec < ds.n_items).all())
    assert bool((torch.sort(rec, dim=1).values[:, 1:] != torch.sort(rec, dim=1).values[:, :-1]).all())  # no repeats
    # no recommended item is a train or val item of that user
    from b200rec import ops
    for split in ("train", "val"):
        p_, i_ = ds.csr(split, device=DEV)
        assert float(ops.hit_matrix(recs[0], 0, p_, i_ if i_.numel() else torch.zeros(1, dtype=torch.int32, device=DEV)).sum()) == 0
    # scores are non-increasing along K and consistent with predict()
    m.eval()
    users = torch.arange(0, 64, device=DEV)
    sc = m.predict(users)
    ids, vals = m.recommend(users, 20)
    assert bool((vals[:, 1:] <= vals[:, :-1]).all())
    assert torch.equal(torch.gather(sc, 1, ids.long()), vals)
    assert torch.equal(vals[:, 0], sc.max(dim=1).values)


def test_c4_spmm_adjoint_and_partition_of_rows():
    """one layer at the 100 M-edge shape: <A x, y> == <x, A y>, and computing the rows in two halves gives the same bits"""
    from b200rec import graph, ops, synth
    g = synth.generate_named("c4", seed=0, device=DEV, heldout=False)
    rows = torch.repeat_interleave(torch.arange(g.n_users, device=DEV), g.train_indptr[1:] - g.train_indptr[:-1])
    a = graph.build_norm_adj(g.n_users, g.n_items, rows, g.train_items, DEV, d=128)
    del rows
    n = a.n_rows
    assert a.nnz == 200_000_000 and a.n_long > 0
    gen = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn((n, 128), device=DEV, generator=gen)
    y = torch.randn((n, 128), device=DEV, generator=gen)
    ax, ay = torch.empty_like(x), torch.empty_like(x)
    ops.spmm(a, x, y=ax)
    ops.spmm(a, y, y=ay)
    lhs, rhs = _dot(ax, y), _dot(x, ay)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs) + 1e-3
    half = torch.full_like(ax, float("nan"))
    ops.spmm(a.row_slice(0, n // 2), x, y=half)
    ops.spmm(a.row_slice(n // 2, n), x, y=half)
    assert torch.equal(half, ax)


def test_c3_inductive_protocol():
    """BASELINE config 3: IGCN on the old 80 % of an Amazon-book-shaped graph, then new users / items represented through
    template aggregation only (generate_feat(is_updating=True)) and the six-way inductive_eval (trainer.py:212-253).
    Size-independent checks: the old/new user split is a partition (user-weighted mean of the two recalls is the overall
    recall), banned item ranges never appear, and the tensor-core and exact evaluation paths agree."""
    import dataset as D
    import model as M
    import trainer as T
    from b200rec import synth
    full = synth.generate_named("c3", seed=0, device=DEV)
    old, n_old_u, n_old_i = synth.inductive_split(full, 0.8, 0.8)
    ds_old = D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": old})
    ds_new = D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": full})
    torch.manual_seed(0)
    m = M.get_model({"name": "IGCN", "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1.0,
                     "device": DEV}, ds_old)
    cfg = {"name": "IGCNTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 0.0, "aux_reg": 0.01, "device": DEV,
           "n_epochs": 1, "batch_size": 2048, "dataloader_num_workers": 0, "test_batch_size": 512, "topks": [1, 5, 10, 15, 20]}
    tr = T.get_trainer(cfg, ds_old, m)
    m.train()
    eng = tr._engine()
    for _ in range(20):
        eng.step()
    eng.sync_optimizer_state()
    assert np.isfinite(eng.meter_avg())
    t_rows = m.embedding.weight.shape[0]
    # re-point at the enlarged dataset: new nodes get rows of F over the OLD template columns only
    m.config["dataset"] = ds_new
    m.n_users, m.n_items = ds_new.n_users, ds_new.n_items
    m.norm_adj = m.generate_graph(ds_new)
    m.feat_mat, _, _, m.row_sum = m.generate_feat(ds_new, is_updating=True)
    m.update_feat_mat()
    assert m.feat_mat.n_cols == t_rows and m.feat_mat.n_rows == ds_new.n_users + ds_new.n_items
    tr2 = T.get_trainer(cfg, ds_new, m)
    res = tr2.inductive_eval(n_old_u, n_old_i)
    te_ptr = full.test_indptr.cpu().numpy()
    has = np.diff(te_ptr) > 0
    w_old, w_new = has[:n_old_u].sum(), has[n_old_u:].sum()
    for k in (5, 20):
        mix = (w_old * res["old_all"]["Recall"][k] + w_new * res["new_all"]["Recall"][k]) / (w_old + w_new)
        assert abs(mix - res["all_all"]["Recall"][k]) < 1e-5
    rec_old_items = tr2.recommend_all("test", banned_items=np.arange(n_old_i, ds_new.n_items))
    assert int(rec_old_items.max()) < n_old_i                              # banned (new) items never recommended
    rec_new_items = tr2.recommend_all("test", banned_items=np.arange(n_old_i))
    assert int(rec_new_items[rec_new_items >= 0].min()) >= n_old_i
    tr2.eval_precision = 0
    assert torch.equal(tr2.recommend_all("test", banned_items=np.arange(n_old_i)), rec_new_items)  # exact == tcgen05


@pytest.mark.parametrize("name", ["SGL", "HALF"])
def test_c2_contrastive_step_full_size(name):
    """SGL / HALF at the headline shape (batch 2048, D 64): the graphed step agrees with the autograd path on the same
    batch, the views keep exactly int(E * aug_rate) pairs, and a redraw re-captures the step."""
    import dataset as D
    import model as M
    import trainer as T
    from b200rec import synth
    g = synth.generate_named("c2", seed=0, device=DEV)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": g})
    cfg = {"name": name + "Trainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4, "contrastive_reg": 0.1, "device": DEV,
           "n_epochs": 1, "batch_size": 2048, "dataloader_num_workers": 0, "test_batch_size": 512, "topks": [1, 5, 10, 15, 20]}
    torch.manual_seed(1)
    m = M.get_model({"name": name, "embedding_size": 64, "n_layers": 3, "aug_rate": 0.8, "device": DEV}, ds)
    n_edges = len(ds)
    assert m.norm_aug_adj1.nnz == 2 * int(n_edges * 0.8)
    tr = T.get_trainer(cfg, ds, m)
    m.train()
    eng = tr._engine()
    eng.step()                                                   # device-sampled batch
    torch.cuda.synchronize()
    batch = eng.batch.clone()
    loss_fused = eng.last_loss()
    w_fused = m.embedding.weight.detach().clone()
    # the same step through autograd + torch.optim.Adam from the same start
    torch.manual_seed(1)
    m2 = M.get_model({"name": name, "embedding_size": 64, "n_layers": 3, "aug_rate": 0.8, "device": DEV}, ds)
    m2.norm_aug_adj1, m2.norm_aug_adj2 = m.norm_aug_adj1, m.norm_aug_adj2
    tr2 = T.get_trainer(dict(cfg, fused=False), ds, m2)
    m2.train()
    loss_ref = tr2._autograd_step(batch, None)
    assert abs(loss_fused - loss_ref) < 5e-6 * max(1.0, abs(loss_ref))
    diff = (w_fused - m2.embedding.weight.detach()).abs()
    assert float(diff.max()) <= 2.5e-4 and float((diff > 1e-5).float().mean()) < 0.02   # Adam's first step, see _assert_adam_close
    v0 = m.aug_version
    m.update_aug_adj()
    assert m.aug_version == v0 + 1 and m.norm_aug_adj1.nnz == 2 * int(n_edges * 0.8)
    for _ in range(3):
        eng.step()                                               # new operands -> the step graph is captured again
    assert np.isfinite(eng.last_loss())
