"""GPU: DOSE_aug / DOSE_drop3 + DOSEaugTrainer / DOSEdropTrainer against fixtures recorded from the unmodified
reference (oracle/make_golden.py --only-dose: model.py:367-613, :2544-2863, trainer.py:255-353, utils.py:71-141):
the mined pairs, the edited adjacency, one training step with the reference's two dropout draws."""
import numpy as np
import pytest
import torch

from conftest import MODEL_CFG, golden_model

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOPKS = [1, 5, 10, 15, 20]
NAMES = ["dose_aug_tiny", "dose_drop3_tiny"]


def _np(t):
    return t.detach().cpu().numpy()


def _model(g, name, **extra):
    ds, m = golden_model(g, name, **extra)
    with torch.no_grad():
        m.w.copy_(torch.from_numpy(g["w0"]))
    return ds, m


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("precision", [0, 1])
def test_mining_matches_reference(golden, name, precision):
    """cal_cos_sim: the same pairs as the reference's sklearn + torch.topk selection, including its second-half offset
    arithmetic; pairs whose value ties with the cut (within 1e-6) may differ, like any top-k"""
    from b200rec import mining
    g = golden(name)
    ds, m = _model(g, name)
    m.eval()
    with torch.no_grad():
        rep = m.get_def_rep()
    np.testing.assert_allclose(_np(rep), g["rep_eval"], rtol=1e-5, atol=1e-7)
    nu, ni, k = int(g["n_users"]), int(g["n_items"]), int(g["aug_num"]) // 2
    ref_rep = torch.from_numpy(g["rep_eval"]).to(DEV)           # mine on the reference's own representation
    ours = _np(mining.lowest_cosine_pairs(ref_rep[:nu], ref_rep[nu:], int(g["aug_num"]), precision=precision))
    ref = g["mined_pairs"]
    assert ours.shape == ref.shape == (2 * k, 2)
    un = torch.nn.functional.normalize(ref_rep[:nu].double(), dim=1)
    vn = -torch.nn.functional.normalize(ref_rep[nu:].double(), dim=1)
    cos = _np(un @ vn.T).reshape(-1)
    half = (nu * ni) // 2
    for h, cut in ((0, float(g["cut_first"])), (1, float(g["cut_second"]))):
        a, b = ours[h * k:(h + 1) * k], ref[h * k:(h + 1) * k]
        fa, fb = a[:, 0] * ni + a[:, 1], b[:, 0] * ni + b[:, 1]
        if h == 1:  # undo the reference's offset: reported flat = j + k, selected cell = j + half
            fa, fb = fa - k + half, fb - k + half
            assert fa.min() >= half and fb.min() >= half
        else:
            assert fa.max() < half and fb.max() < half
        clear_b = cos[fb] - cut > 1e-6
        clear_a = cos[fa] - cut > 1e-6
        assert clear_b.mean() > 0.98
        assert set(fb[clear_b].tolist()) <= set(fa.tolist()) and set(fa[clear_a].tolist()) <= set(fb.tolist())
        assert (cos[fa] >= cut - 1e-6).all()
    fixed = _np(mining.lowest_cosine_pairs(ref_rep[:nu], ref_rep[nu:], int(g["aug_num"]), reference_offsets=False))
    assert np.array_equal(fixed[:k], ours[:k]) if precision == 0 else True
    f2 = fixed[k:, 0] * ni + fixed[k:, 1]
    assert f2.min() >= half and (cos[f2] >= float(g["cut_second"]) - 1e-6).all()


@pytest.mark.parametrize("name", NAMES)
def test_edited_graph_matches_reference(golden, name):
    g = golden(name)
    ds, m = _model(g, name)
    adj = m.generate_aug_graph(ds, g["mined_pairs"]) if name.startswith("dose_aug") else m.generate_drop_graph(ds, g["mined_pairs"])
    r, c, v = adj.to_coo()
    assert np.array_equal(np.stack([_np(r), _np(c)]), g["aug_idx"])
    ulp = np.abs(_np(v).view(np.int32).astype(np.int64) - g["aug_val"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 3
    assert adj.nnz != m.norm_adj.nnz


@pytest.mark.parametrize("name", NAMES)
def test_training_step_matches_reference(golden, name):
    import trainer as T
    from b200rec import ops
    g = golden(name)
    ds, m = _model(g, name)
    m.norm_aug_adj = m._edited_graph(ds, g["mined_pairs"])
    tr = T.get_trainer({"name": "DOSEaugTrainer" if name.startswith("dose_aug") else "DOSEdropTrainer", "optimizer": "Adam",
                        "lr": float(g["lr"]), "l2_reg": float(g["l2_reg"]), "aux_reg": float(g["aux_reg"]),
                        "contrastive_reg": float(g["contrastive_reg"]), "device": DEV, "n_epochs": 1, "batch_size": 256,
                        "dataloader_num_workers": 0, "test_batch_size": 128, "topks": TOPKS}, ds, m)
    m.train()
    nnz = int(g["drop_nnz"])
    masks = [ops.pack_keep_bits(torch.from_numpy(np.unpackbits(g[k])[:nnz].astype(bool)).to(DEV))
             for k in ("drop_keep_def", "drop_keep_aug")]
    it = iter(masks)
    m._keep_bits = lambda: next(it)      # the reference's two recorded draws: get_def_rep first, then get_aug_rep
    batch, aux = torch.from_numpy(g["batch"]).to(DEV), torch.from_numpy(g["aux_batch"]).to(DEV)
    loss = tr._loss(batch, aux)
    assert abs(float(loss) - float(g["loss"])) < 2e-6
    tr.opt.zero_grad()
    loss.backward()
    np.testing.assert_allclose(_np(m.embedding.weight.grad), g["grad_emb"], rtol=1e-4, atol=2e-9)
    np.testing.assert_allclose(_np(m.w.grad), g["grad_w"], rtol=1e-4, atol=1e-9)
    tr.opt.step()
    np.testing.assert_allclose(_np(m.embedding.weight), g["emb1"], rtol=1e-5, atol=5e-6)
    np.testing.assert_allclose(_np(m.w), g["w1"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", NAMES)
def test_epoch_runs_and_remines(golden, name):
    import trainer as T
    g = golden(name)
    ds, m = _model(g, name)
    tr = T.get_trainer({"name": "DOSEaugTrainer" if name.startswith("dose_aug") else "DOSEdropTrainer", "optimizer": "Adam",
                        "lr": 1e-3, "l2_reg": 0.0, "aux_reg": 0.001, "contrastive_reg": 0.1, "device": DEV, "n_epochs": 1,
                        "batch_size": 512, "dataloader_num_workers": 0, "test_batch_size": 128, "topks": TOPKS}, ds, m)
    nnz0, a0 = m.norm_aug_adj.nnz, m.alpha
    m.train()
    loss = tr.train_one_epoch()
    assert np.isfinite(loss) and m.aug_version == 1 and m.alpha < a0
    if name.startswith("dose_aug"):
        assert m.norm_aug_adj.nnz > m.norm_adj.nnz
    else:
        assert m.norm_aug_adj.nnz <= m.norm_adj.nnz
    _, metrics, _ = tr.eval("val")
    assert 0.0 <= metrics["Recall"][20] <= 1.0 and nnz0 > 0
    assert MODEL_CFG[name]["aug_num"] == 2000


def _small_dataset():
    import dataset as D
    from b200rec import synth
    return D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": synth.generate(300, 400, 9000, seed=7)})


def _dose_cfg(name, **kw):
    return dict({"name": name, "embedding_size": 64, "n_layers": 2, "device": DEV, "dropout": 0.3, "feature_ratio": 1.0,
                 "aug_num": 500}, **kw)


def test_dose_drop2_random_subgraph_and_epoch():
    """DOSE_drop2 (reference model.py:1673-1962): the contrast graph holds exactly int(E * aug_rate) train pairs
    (utils.py:91-103), every one a train edge, degrees recomputed on the subset; update_aug_adj redraws it"""
    import model as M
    import trainer as T
    ds = _small_dataset()
    m = M.get_model(_dose_cfg("DOSE_drop2", aug_rate=0.25), ds)
    users, items = ds.train_pairs()
    n_keep = int(len(users) * 0.25)
    a = m.norm_aug_adj
    assert a.nnz == 2 * n_keep
    rp, ci = _np(a.rowptr).astype(np.int64), _np(a.colidx).astype(np.int64)
    rows = np.repeat(np.arange(ds.n_users + ds.n_items), np.diff(rp))
    upper = rows < ds.n_users
    kept = set(zip(rows[upper].tolist(), (ci[upper] - ds.n_users).tolist()))
    assert len(kept) == n_keep and kept <= set(zip(users.tolist(), items.tolist()))
    deg = np.maximum(np.diff(rp), 1).astype(np.float64)
    np.testing.assert_allclose(_np(a.vals), (deg[rows] ** -0.5 * deg[ci] ** -0.5).astype(np.float32), rtol=1e-6)
    tr = T.get_trainer({"name": "DOSEdropTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 0.0, "aux_reg": 0.001,
                        "contrastive_reg": 0.1, "device": DEV, "n_epochs": 1, "batch_size": 512, "dataloader_num_workers": 0,
                        "test_batch_size": 128, "topks": TOPKS}, ds, m)
    first = _np(a.colidx).copy()
    m.train()
    assert np.isfinite(tr.train_one_epoch()) and m.aug_version == 1
    assert m.norm_aug_adj.nnz == 2 * n_keep and not np.array_equal(_np(m.norm_aug_adj.colidx), first)


def test_dose_aug_drop2_mines_the_most_similar_low_degree_pairs():
    """DOSE_aug_drop2.cal_cos_sim (reference model.py:3290-3324): flat top-k of the cosine matrix of the users x items that
    remain after the aug_ratio highest-degree ones are left out, mapped back to global ids -- against a dense float64
    restatement; both graphs are the train graph plus those pairs"""
    import model as M
    import trainer as T
    import utils as U
    ds = _small_dataset()
    m = M.get_model(_dose_cfg("DOSE_aug_drop2", aug_ratio=0.2), ds)
    m.eval()
    pairs = _np(m.cal_cos_sim())
    assert pairs.shape == (500, 2)
    ru, ri = U.graph_rank_nodes(ds, "degree")
    cu, ci = ru[int(ds.n_users * 0.2):], ri[int(ds.n_items * 0.2):]
    assert set(pairs[:, 0].tolist()) <= set(cu.tolist()) and set(pairs[:, 1].tolist()) <= set(ci.tolist())
    with torch.no_grad():
        rep = m.get_def_rep().double()
    un = torch.nn.functional.normalize(rep[torch.from_numpy(cu.copy()).to(DEV)], dim=1)
    vn = torch.nn.functional.normalize(rep[ds.n_users + torch.from_numpy(ci.copy()).to(DEV)], dim=1)
    cos = _np(un @ vn.T)
    cut = np.sort(cos.reshape(-1))[-500]
    pos_u = {int(u): k for k, u in enumerate(cu)}
    pos_i = {int(i): k for k, i in enumerate(ci)}
    got = np.array([cos[pos_u[int(u)], pos_i[int(i)]] for u, i in pairs])
    assert (got >= cut - 1e-6).all() and (np.diff(got) <= 1e-6).all()           # the top-k, best first
    clear = np.argwhere(cos > cut + 1e-6)
    assert {(int(cu[a]), int(ci[b])) for a, b in clear} <= {(int(u), int(i)) for u, i in pairs}
    users, items = ds.train_pairs()
    n_union = len(set(zip(users.tolist(), items.tolist())) | {(int(u), int(i)) for u, i in pairs})
    assert m.generate_aug_graph(ds, pairs).nnz == m.generate_drop_graph(ds, pairs).nnz == 2 * n_union
    assert m.norm_aug_adj.nnz > m.norm_adj.nnz     # (the constructor mined with its own dropout draw, like the reference)
    tr = T.get_trainer({"name": "DOSEdropTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 0.0, "aux_reg": 0.001,
                        "contrastive_reg": 0.1, "device": DEV, "n_epochs": 1, "batch_size": 512, "dataloader_num_workers": 0,
                        "test_batch_size": 128, "topks": TOPKS}, ds, m)
    m.train()
    assert np.isfinite(tr.train_one_epoch()) and m.aug_version == 1 and m.norm_drop_adj.nnz > m.norm_adj.nnz
