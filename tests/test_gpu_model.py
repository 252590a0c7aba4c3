"""GPU parity tests, drop-in surface: model.py / trainer.py of this repo against fixtures produced by the unmodified
reference (tests/golden, oracle/make_golden.py).  Same inputs (graph, initial weights, recorded batch and dropout
draw) -> same representations, loss, gradients, post-Adam weights (fp32, 1e-5 relative) and the same top-K ids and
metrics."""
import numpy as np
import pytest
import torch

from conftest import MODEL_CFG, golden_maps, golden_model

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOPKS = [1, 5, 10, 15, 20]
ALL = ["lightgcn_tiny", "lightgcn_d128", "mf_tiny", "igcn_tiny", "igcn_fr_tiny", "imf_tiny"]


def _np(t):
    return t.detach().cpu().numpy()


def _trainer(g, name, ds, m, **extra):
    import trainer as T
    is_igcn = MODEL_CFG[name]["name"] in ("IGCN", "IMF")
    cfg = {"name": "IGCNTrainer" if is_igcn else "BPRTrainer", "optimizer": "Adam", "lr": float(g["lr"]),
           "l2_reg": float(g["l2_reg"]), "aux_reg": float(g["aux_reg"]), "device": DEV, "n_epochs": 1,
           "batch_size": int(g["batch"].shape[0]), "dataloader_num_workers": 0, "test_batch_size": 128, "topks": TOPKS}
    cfg.update(extra)
    return T.get_trainer(cfg, ds, m)


def _keep_bits(g):
    from b200rec import ops
    keep = np.unpackbits(g["drop_keep"])[: int(g["drop_nnz"])].astype(bool)
    return ops.pack_keep_bits(torch.from_numpy(keep).to(DEV))


@pytest.mark.parametrize("name", ["igcn_tiny", "igcn_fr_tiny", "imf_tiny"])
def test_template_features(golden, name):
    g = golden(name)
    ds, m = golden_model(g, name)
    assert (m.user_map, m.item_map) == golden_maps(g)                          # template choice (graph_rank_nodes)
    r, c, _ = m.feat_mat.fwd.to_coo()
    assert np.array_equal(np.stack([_np(r), _np(c)]), g["feat_idx"])           # F structure: bit-exact
    assert np.array_equal(_np(m.row_sum), g["row_sum"])
    np.testing.assert_allclose(_np(m.row_scale)[_np(r)], g["feat_val_a1"], rtol=1e-6)
    m.feat_mat_anneal()
    assert abs(m.alpha - float(g["alpha_after"])) < 1e-12
    np.testing.assert_allclose(_np(m.row_scale)[_np(r)], g["feat_val_anneal"], rtol=1e-6)
    # transposed operand is consistent with the forward one
    tr, tc = _np(m.feat_mat.bwd.to_coo()[0]), _np(m.feat_mat.bwd.to_coo()[1])
    eid = _np(m.feat_mat.bwd.eid)
    assert np.array_equal(_np(r)[eid], tc) and np.array_equal(_np(c)[eid], tr)


@pytest.mark.parametrize("name", [n for n in ALL if not n.startswith("mf")])
def test_get_rep_eval(golden, name):
    g = golden(name)
    ds, m = golden_model(g, name)
    m.eval()
    with torch.no_grad():
        rep = m.get_rep()
        assert m.get_rep() is rep      # eval-mode cache
    np.testing.assert_allclose(_np(rep), g["rep_eval"], rtol=1e-5, atol=1e-7)
    sc = m.predict(torch.arange(8, device=DEV))
    np.testing.assert_allclose(_np(sc), g["scores_head"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ALL)
def test_bpr_forward_autograd_path(golden, name):
    """the reference-shaped loop: bpr_forward -> torch loss -> backward -> torch.optim.Adam"""
    g = golden(name)
    ds, m = golden_model(g, name, dropout_rng="host")
    tr = _trainer(g, name, ds, m, fused=False)
    m.train()
    batch = torch.from_numpy(g["batch"]).to(DEV)
    is_igcn = "drop_keep" in g
    if is_igcn:
        bits = _keep_bits(g)
        m._keep_bits = lambda: bits          # inject the reference's recorded torch.rand draw
    ur, pr, nr, l2 = m.bpr_forward(batch[:, 0].contiguous(), batch[:, 1].contiguous(), batch[:, 2].contiguous())
    np.testing.assert_allclose(_np(ur), g["users_r"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(_np(pr), g["pos_r"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(_np(nr), g["neg_r"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(_np(l2), g["l2_norm_sq"], rtol=1e-5)
    aux = torch.from_numpy(g["aux_batch"]).to(DEV) if "aux_batch" in g else None
    loss = tr._loss(batch, aux)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    tr.opt.zero_grad()
    loss.backward()
    if name.startswith("mf"):
        np.testing.assert_allclose(_np(m.user_embedding.weight.grad), g["grad_user"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(_np(m.item_embedding.weight.grad), g["grad_item"], rtol=1e-4, atol=1e-9)
    else:
        np.testing.assert_allclose(_np(m.embedding.weight.grad), g["grad_emb"], rtol=1e-4, atol=1e-9)
    if "grad_w" in g:
        np.testing.assert_allclose(_np(m.w.grad), g["grad_w"], rtol=1e-4, atol=1e-9)
    tr.opt.step()
    if name.startswith("mf"):
        np.testing.assert_allclose(_np(m.user_embedding.weight), g["user_emb1"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(_np(m.item_embedding.weight), g["item_emb1"], rtol=1e-5, atol=2e-6)
    else:
        np.testing.assert_allclose(_np(m.embedding.weight), g["emb1"], rtol=1e-5, atol=5e-6)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_engine_step(golden, name, use_graph):
    """the CUDA-graphed fused step fed the reference's recorded batch: loss, gradient and post-Adam weights.
    (Adam's first update is lr*g/(|g|+eps): where |g| ~ eps=1e-8 it amplifies the fp32 rounding of g, hence the
    5e-6 absolute slack on a 1e-3 step.)"""
    g = golden(name)
    ds, m = golden_model(g, name)
    tr = _trainer(g, name, ds, m)
    m.train()
    eng = tr._engine()
    eng.use_graph = use_graph
    hb = torch.from_numpy(g["batch"]).pin_memory()
    ha = torch.from_numpy(g["aux_batch"]).pin_memory() if "aux_batch" in g else None
    hk = _keep_bits(g) if "drop_keep" in g else None
    eng.step(host_batch=hb, host_aux_batch=ha, host_keep_bits=hk)
    assert abs(eng.last_loss() - float(g["loss"])) < 2e-6
    assert abs(eng.meter_avg() - float(g["loss"])) < 2e-6
    if name.startswith("mf"):
        np.testing.assert_allclose(_np(m.user_embedding.weight.grad), g["grad_user"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(_np(m.user_embedding.weight), g["user_emb1"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(_np(m.item_embedding.weight), g["item_emb1"], rtol=1e-5, atol=2e-6)
    else:
        np.testing.assert_allclose(_np(m.embedding.weight.grad), g["grad_emb"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(_np(m.embedding.weight), g["emb1"], rtol=1e-5, atol=5e-6)
    if "grad_w" in g:
        np.testing.assert_allclose(_np(m.w.grad), g["grad_w"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(_np(m.w), g["w1"], rtol=1e-5, atol=2e-6)
    eng.sync_optimizer_state()
    assert int(tr.opt.state[next(iter(m.parameters()))]["step"]) == 1


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("name", ["lightgcn_tiny", "mf_tiny", "igcn_tiny", "lightgcn_d128"])
def test_eval_topk_and_metrics(golden, name, precision):
    """precision 0 = exact fp32 CUDA-core scoring, 1 = tcgen05 bf16 candidates + exact re-score (same ids by construction)"""
    g = golden(name)
    ds, m = golden_model(g, name)
    tr = _trainer(g, name, ds, m, eval_precision=precision)
    for split in ("train", "val", "test"):
        rec = _np(tr.recommend_all(split))
        ref_ids, ref_val = g["topk_ids_" + split], g["topk_val_" + split]
        sep = np.ones_like(ref_ids, dtype=bool)
        d = np.abs(np.diff(ref_val, axis=1)) > 1e-6
        sep[:, 1:] &= d
        sep[:, :-1] &= d
        assert sep.mean() > 0.99
        assert np.array_equal(rec[sep], ref_ids[sep])
        _, metrics, _ = tr.eval(split)
        for mname in ("Precision", "Recall", "NDCG"):
            ours = np.array([metrics[mname][k] for k in TOPKS])
            np.testing.assert_allclose(ours, g["metric_%s_%s" % (mname, split)], rtol=1e-5, atol=1e-7)
    _, metrics, _ = tr.eval("test", banned_items=np.arange(int(g["banned_lo"]), int(g["banned_hi"])))
    for mname in ("Precision", "Recall", "NDCG"):
        ours = np.array([metrics[mname][k] for k in TOPKS])
        np.testing.assert_allclose(ours, g["metric_%s_test_banned" % mname], rtol=1e-5, atol=1e-7)


def test_train_epoch_device_sampler_runs_and_learns(golden):
    g = golden("lightgcn_tiny")
    ds, m = golden_model(g, "lightgcn_tiny")
    tr = _trainer(g, "lightgcn_tiny", ds, m, lr=5e-3)
    m.train()
    l0 = tr.train_one_epoch()
    for _ in range(5):
        l1 = tr.train_one_epoch()
    assert np.isfinite(l0) and l1 < l0
    _, metrics, _ = tr.eval("val")
    assert 0.0 <= metrics["Recall"][20] <= 1.0
    # host-sampler mode (the reference's DataLoader batches, H2D per step) also runs
    tr2 = _trainer(g, "lightgcn_tiny", ds, m, sampler="host")
    assert np.isfinite(tr2.train_one_epoch())


@pytest.mark.parametrize("name", ["lightgcn_tiny", "igcn_tiny"])
def test_ragged_last_batch_in_fused_epoch(golden, name):
    """len(dataset) % batch_size != 0 with the host sampler: the tail batch runs eagerly inside a fused epoch.  Its Adam
    bias correction must use the epoch's real step count and its loss must enter the epoch average: the epoch
    equals the all-autograd epoch on the same DataLoader batches."""
    import utils
    g = golden(name)
    res = []
    for fused in (True, False):
        ds, m = golden_model(g, name, dropout=0.0) if name.startswith("igcn") else golden_model(g, name)
        tr = _trainer(g, name, ds, m, fused=fused, sampler="host", batch_size=256)
        assert len(ds) % 256 != 0
        m.train()
        losses = []
        for ep in range(2):
            utils.set_seed(77 + ep)      # the reference's host sampler draws from random / np.random
            losses.append(tr.train_one_epoch())
        steps = int(tr.opt.state[m.embedding.weight]["step"])
        res.append((losses, steps, _np(m.embedding.weight).copy()))
    (l_f, s_f, w_f), (l_a, s_a, w_a) = res
    assert s_f == s_a == 2 * -(-len(ds) // 256)
    np.testing.assert_allclose(l_f, l_a, rtol=2e-6)
    np.testing.assert_allclose(w_f, w_a, rtol=1e-4, atol=3e-5)   # a stale bias correction moves the tail step by ~1e-2


def test_checkpoint_roundtrip(golden, tmp_path):
    g = golden("igcn_fr_tiny")
    ds, m = golden_model(g, "igcn_fr_tiny")
    m.feat_mat_anneal()
    p = str(tmp_path / "m.pth")
    m.save(p)
    ck = torch.load(p, weights_only=False)
    assert set(ck) == {"sate_dict", "user_map", "item_map", "alpha"}          # reference checkpoint layout
    assert set(ck["sate_dict"]) == {"embedding.weight", "w"}
    ds2, m2 = golden_model(g, "igcn_fr_tiny")
    with torch.no_grad():
        m2.embedding.weight.zero_()
    m2.load(p)
    m.eval(); m2.eval()
    with torch.no_grad():
        assert torch.equal(m.get_rep(), m2.get_rep())


def test_inductive_eval_matches_reference(golden, capsys):
    """IGCN built on the old nodes, re-pointed at the enlarged dataset (generate_feat(is_updating=True), new adjacency),
    then the six passes of inductive_eval (trainer.py:212-253): the printed result lines equal the reference's."""
    import re
    import dataset as D
    import model as M
    import trainer as T
    from b200rec import synth
    g = golden("igcn_inductive")
    t = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.int64))  # noqa: E731

    def ds_of(prefix, nu, ni):
        graph = synth.SynthGraph(nu, ni, t(prefix + "_train_indptr"), t(prefix + "_train_items"), t(prefix + "_val_indptr"),
                                 t(prefix + "_val_items"), t(prefix + "_test_indptr"), t(prefix + "_test_items"))
        return D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": graph})

    n_old_u, n_old_i = int(g["n_old_users"]), int(g["n_old_items"])
    ds_old = ds_of("old", n_old_u, n_old_i)
    ds_new = ds_of("new", int(g["new_n_users"]), int(g["new_n_items"]))
    m = M.get_model({"name": "IGCN", "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1.0,
                     "device": DEV}, ds_old)
    with torch.no_grad():
        m.embedding.weight.copy_(torch.from_numpy(g["emb0"]))
    # what a driver does to evaluate new users / items (the reference ships none; IGCN.load shows the sequence)
    m.config["dataset"] = ds_new
    m.n_users, m.n_items = ds_new.n_users, ds_new.n_items
    m.norm_adj = m.generate_graph(ds_new)
    m.feat_mat, _, _, m.row_sum = m.generate_feat(ds_new, is_updating=True)
    m.update_feat_mat()
    r, c, _ = m.feat_mat.fwd.to_coo()
    assert np.array_equal(np.stack([_np(r), _np(c)]), g["feat_idx_new"])     # new nodes use template columns only
    m.eval()
    with torch.no_grad():
        np.testing.assert_allclose(_np(m.get_rep()), g["rep_new"], rtol=1e-5, atol=1e-7)
    tr = T.get_trainer({"name": "IGCNTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 0.0, "aux_reg": 0.01,
                        "device": DEV, "dataloader_num_workers": 0, "topks": TOPKS, "n_epochs": 1, "batch_size": 256,
                        "test_batch_size": 128}, ds_new, m)
    capsys.readouterr()
    tr.inductive_eval(n_old_u, n_old_i)
    ours = [l for l in capsys.readouterr().out.splitlines() if "result." in l]
    ref = [str(x) for x in g["inductive_lines"]]
    assert len(ours) == 6
    num = re.compile(r"-?\d+\.\d+")
    for a, b in zip(ours, ref):
        assert a.split(" result.")[0] == b.split(" result.")[0]
        va, vb = [float(x) for x in num.findall(a)], [float(x) for x in num.findall(b)]
        assert len(va) == len(vb) == 15
        np.testing.assert_allclose(va, vb, atol=2e-3)   # printed with 3 decimals
    assert ds_new.test_data is not None and len(ds_new.test_data) == ds_new.n_users  # test_data restored


def test_train_loop_checkpoints_and_early_stop(golden, tmp_path, monkeypatch, capsys):
    """BasicTrainer.train (trainer.py:58-113): eval('train') + eval('val') per epoch, best-NDCG checkpoint rotation in
    ./checkpoints, patience, reload of the best checkpoint at the end."""
    import os
    g = golden("lightgcn_tiny")
    ds, m = golden_model(g, "lightgcn_tiny")
    monkeypatch.chdir(tmp_path)
    tr = _trainer(g, "lightgcn_tiny", ds, m, n_epochs=4, lr=5e-3, max_patience=2)
    best = tr.train(verbose=True)
    out = capsys.readouterr().out
    assert out.count("Validation result.") >= 1 and "Epoch 0/4" in out
    files = os.listdir(tmp_path / "checkpoints")
    assert len(files) == 1 and files[0].startswith("LightGCN_BPRTrainer_SyntheticDataset_") and files[0].endswith(".pth")
    assert abs(float(files[0].rsplit("_", 1)[1][:-4]) - best * 100) < 1e-3        # file name carries NDCG*100 (3 decimals)
    assert tr.save_path == os.path.join("checkpoints", files[0]) and best == tr.best_ndcg > 0
    sd = torch.load(tr.save_path)
    assert list(sd) == ["embedding.weight"] and torch.equal(sd["embedding.weight"].to(DEV), m.embedding.weight.data)


def test_degenerate_graph_rows_and_k_larger_than_catalogue():
    """users without train items, items nobody interacted with (degree clamped to 1, model.py:92), K > unmasked items"""
    import dataset as D
    import model as M
    import trainer as T
    from b200rec import synth
    tp = torch.tensor([0, 0, 3, 3, 4, 6], dtype=torch.int64)          # users 0 and 2 have no train items
    ti = torch.tensor([1, 4, 7, 0, 2, 7], dtype=torch.int64)          # items 3, 5, 6, 8 are never used
    e = torch.zeros(6, dtype=torch.int64)
    graph = synth.SynthGraph(5, 9, tp, ti, e, ti[:0], tp, ti)          # test = train (just to have eval rows)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": graph})
    m = M.get_model({"name": "LightGCN", "embedding_size": 64, "n_layers": 2, "device": DEV}, ds)
    dinv = _np(m.norm_adj.dinv)
    assert dinv[0] == 1.0 and dinv[5 + 3] == 1.0                        # isolated nodes: deg = max(1, 0)
    tr = T.get_trainer({"name": "BPRTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4, "device": DEV, "n_epochs": 1,
                        "batch_size": 64, "dataloader_num_workers": 0, "test_batch_size": 4, "topks": TOPKS}, ds, m)
    m.train()
    assert np.isfinite(tr.train_one_epoch())
    assert set(_np(tr._engine().batch[:, 0]).tolist()) <= {1, 3, 4}     # the sampler never draws a user with an empty row
    rec = _np(tr.recommend_all("val"))                                  # K = 20 > 9 items: padded with -1
    assert rec.shape == (5, 20) and (rec[:, 9:] == -1).all()
    assert sorted(rec[1, :6].tolist()) == [0, 2, 3, 5, 6, 8]            # user 1's train items 1, 4, 7 are masked
    _, metrics, _ = tr.eval("test")
    assert 0.0 <= metrics["Recall"][20] <= 1.0


def test_arbitrary_banned_items_and_nonfinite_guard(golden):
    """trainer.eval(banned_items=<any index array>) (trainer.py:166-167) == dense scores with those columns at -inf;
    a contiguous range takes the in-kernel path, anything else scores against the compacted item table.  A diverged model
    (NaN in the representation) raises instead of reporting all-zero metrics."""
    g = golden("lightgcn_tiny")
    ds, m = golden_model(g, "lightgcn_tiny")
    tr = _trainer(g, "lightgcn_tiny", ds, m)
    rng = np.random.default_rng(4)
    banned = np.unique(rng.choice(ds.n_items, size=ds.n_items // 3, replace=False))
    assert banned.max() - banned.min() + 1 != banned.size
    rec = _np(tr.recommend_all("test", banned_items=banned))
    m.eval()
    with torch.no_grad():
        sc = _np(m.predict(torch.arange(ds.n_users, device=DEV))).astype(np.float64)
    for which in ("train", "val"):
        ptr, idx = (a.cpu().numpy() for a in ds.csr(which))
        for u in range(ds.n_users):
            sc[u, idx[ptr[u]:ptr[u + 1]]] = -np.inf
    sc[:, banned] = -np.inf
    order = np.argsort(-sc, axis=1, kind="stable")[:, :20]
    vals = np.take_along_axis(sc, order, axis=1)
    sep = np.ones_like(order, dtype=bool)
    d = np.abs(np.diff(vals, axis=1)) > 1e-6
    sep[:, 1:] &= d
    sep[:, :-1] &= d
    assert sep.mean() > 0.99 and np.array_equal(rec[sep], order[sep])
    assert not np.isin(rec, banned).any()
    # same metrics through eval(), and the contiguous special case agrees with the general path
    lo, hi = 100, 260
    r1 = _np(tr.recommend_all("test", banned_items=np.arange(lo, hi)))
    r2 = _np(tr.recommend_all("test", banned_items=np.concatenate([np.arange(lo, hi), [lo]])[::-1]))   # duplicates, unordered
    assert np.array_equal(r1, r2)
    with torch.no_grad():
        m.embedding.weight[3, 5] = float("nan")
    m._rep_cache = None
    with pytest.raises(FloatingPointError):
        tr.eval("val")
