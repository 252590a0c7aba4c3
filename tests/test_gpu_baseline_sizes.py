"""GPU parity at BASELINE.json sizes (VERDICT r1 "parity is green only on toys").

  * C1 (29 858 x 40 981, 1 027 370 edges, LightGCN L=3 D=64, batch 2048) against `tests/golden/c1_ref.npz`, produced by
    the UNMODIFIED reference (oracle/make_golden.py --only-c1): normalised adjacency, eval representation, one training
    step (loss / gradient / post-Adam weights) and eval('val'|'test') with Recall/NDCG@20
    (/root/reference/trainer.py:115-210, :412-429; model.py:89-127).
  * C2 (LightGCN) and C3 (IGCN, inductive template layer + edge dropout + auxiliary loss) LIVE against the CPU port
    (oracle/ref_port.py, itself pinned to the reference fixtures): one engine step on the device-sampled batch and the
    top-20 of 1 024 users.
  * C5 (2M x 1M x 128 sweep): tcgen05 path == exact path on a 4 096-user slice with the real train CSR masked.
Same bars as the tiny fixtures: loss 2e-6, gradients 1e-4 rel, post-Adam 1e-5 rel + 5e-6 abs, representations 1e-5 rel,
metrics 1e-5, ids equal wherever the reference's adjacent scores differ by more than 1e-6."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOPKS = [1, 5, 10, 15, 20]


def _np(t):
    return t.detach().cpu().numpy()


def _build(shape, model_cfg, trainer_cfg, graph=None, emb0=None):
    import dataset as D
    import model as M
    import trainer as T
    from b200rec import synth
    g = graph if graph is not None else synth.generate_named(shape, seed=0, device=DEV)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": g})
    torch.manual_seed(2021)
    m = M.get_model(dict(model_cfg, device=DEV), ds)
    if emb0 is not None:
        with torch.no_grad():
            m.embedding.weight.copy_(torch.from_numpy(emb0))
    cfg = {"optimizer": "Adam", "lr": 1e-3, "device": DEV, "n_epochs": 1, "batch_size": 2048, "dataloader_num_workers": 0,
           "test_batch_size": 512, "topks": TOPKS}
    cfg.update(trainer_cfg)
    return g, ds, m, T.get_trainer(cfg, ds, m)


@pytest.fixture(scope="module")
def c1():
    from b200rec import synth
    f = load_golden("c1_ref")
    graph = synth.generate_named("c1", seed=0, device=DEV)
    assert synth.fingerprint(graph) == int(f["fingerprint"]), "the generator no longer draws the fixture's graph"
    emb0 = synth.hashed_embedding(graph, 64)
    assert abs(float(np.abs(emb0.astype(np.float64)).sum()) - float(f["emb0_checksum"])) < 1e-6
    g, ds, m, tr = _build("c1", {"name": "LightGCN", "embedding_size": 64, "n_layers": 3},
                          {"name": "BPRTrainer", "l2_reg": float(f["l2_reg"])}, graph=graph, emb0=emb0)
    return f, ds, m, tr, emb0


def test_c1_adjacency_matches_reference(c1):
    f, ds, m, tr, _ = c1
    a = m.norm_adj
    assert a.nnz == int(f["adj_nnz"])
    r, c, v = a.to_coo()
    pos = torch.from_numpy(f["adj_pos"]).to(DEV)
    assert np.array_equal(np.stack([_np(r[pos]), _np(c[pos])]), f["adj_idx_at"])          # coalesced order: bit-exact
    ours, ref = _np(v[pos]), f["adj_val_at"]
    ulp = np.abs(ours.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert ulp.max() <= 3                                                                 # numpy's power(deg,-0.5) is itself 1 ulp off


def test_c1_representation_matches_reference(c1):
    f, ds, m, tr, _ = c1
    m.eval()
    with torch.no_grad():
        rep = m.get_rep()
    rows = torch.from_numpy(f["rows_kept"]).to(DEV)
    np.testing.assert_allclose(_np(rep[rows]), f["rep_eval_rows"], rtol=1e-5, atol=1e-7)
    assert abs(float(rep.double().abs().sum()) - float(f["rep_eval_abs_sum"])) < 1e-6 * float(f["rep_eval_abs_sum"])


@pytest.mark.parametrize("precision", [1, 0])
def test_c1_eval_metrics_and_topk_match_reference(c1, precision):
    f, ds, m, tr, emb0 = c1
    with torch.no_grad():
        m.embedding.weight.copy_(torch.from_numpy(emb0))
    m._rep_cache = None
    tr.eval_precision = precision
    for split in ("val", "test"):
        _, metrics, _ = tr.eval(split)
        for name in ("Precision", "Recall", "NDCG"):
            ours = np.array([metrics[name][k] for k in TOPKS])
            np.testing.assert_allclose(ours, f["metric_%s_%s" % (name, split)], rtol=1e-5, atol=1e-7)
    rec = _np(tr.recommend_all("test"))[: f["topk_ids_test"].shape[0]]
    ref_ids, ref_val = f["topk_ids_test"], f["topk_val_test"]
    d = np.abs(np.diff(ref_val, axis=1)) > 1e-6
    sep = np.ones_like(ref_ids, dtype=bool)
    sep[:, 1:] &= d
    sep[:, :-1] &= d
    assert sep.mean() > 0.97
    assert np.array_equal(rec[sep], ref_ids[sep])


@pytest.mark.parametrize("use_graph", [True, False])
def test_c1_training_step_matches_reference(c1, use_graph):
    f, ds, m, tr, emb0 = c1
    import trainer as T
    with torch.no_grad():
        m.embedding.weight.copy_(torch.from_numpy(emb0))
    tr2 = T.get_trainer(dict(tr.config, name="BPRTrainer"), ds, m)   # fresh optimiser state
    m.train()
    eng = tr2._engine()
    eng.use_graph = use_graph
    eng.step(host_batch=torch.from_numpy(f["batch"]).pin_memory())
    assert abs(eng.last_loss() - float(f["loss"])) < 2e-6
    rows = torch.from_numpy(f["rows_kept"]).to(DEV)
    g = m.embedding.weight.grad
    np.testing.assert_allclose(_np(g[rows]), f["grad_rows"], rtol=1e-4, atol=1e-9)
    assert abs(float(g.double().abs().sum()) - float(f["grad_abs_sum"])) < 1e-5 * float(f["grad_abs_sum"])
    np.testing.assert_allclose(_np(m.embedding.weight[rows]), f["emb1_rows"], rtol=1e-5, atol=5e-6)
    with torch.no_grad():
        m.embedding.weight.copy_(torch.from_numpy(emb0))
    m._rep_cache = None


# ------------------------------------------------------------------------------------------- live against the CPU port
def _port_topk(port, graph, n_users, split="test"):
    from oracle import ref_port as rp
    port.eval()
    with torch.no_grad():
        scores = port.predict(torch.arange(n_users, dtype=torch.int64)).numpy()
    for which in (("train", "val") if split == "test" else ("train",)):
        ptr = _np(getattr(graph, which + "_indptr"))[: n_users + 1]
        idx = _np(getattr(graph, which + "_items"))[: int(ptr[-1])]
        for u in range(n_users):
            scores[u, idx[ptr[u]:ptr[u + 1]]] = -np.inf
    return rp.topk_tiebreak(scores, 20)


def _check_topk(rec, ids, vals):
    d = np.abs(np.diff(vals, axis=1)) > 1e-6
    sep = np.ones_like(ids, dtype=bool)
    sep[:, 1:] &= d
    sep[:, :-1] &= d
    assert sep.mean() > 0.97
    assert np.array_equal(rec[sep], ids[sep])


def test_c2_step_and_topk_against_port():
    from oracle import ref_port as rp
    g, ds, m, tr = _build("c2", {"name": "LightGCN", "embedding_size": 64, "n_layers": 3},
                          {"name": "BPRTrainer", "l2_reg": 1e-4})
    emb0 = _np(m.embedding.weight).copy()
    users, items = ds.train_pairs()
    port = rp.LightGCNPort(ds.n_users, ds.n_items, users, items, emb0, 3)
    ids, vals = _port_topk(port, g, 1024)
    rec = _np(tr.recommend_all("test"))[:1024]
    _check_topk(rec, ids, vals)
    # one optimiser step on the batch the device sampler draws (the oracle's C restatement of the stream drives the port)
    m.train()
    eng = tr._engine()
    eng.step()
    batch = eng.batch.cpu()
    opt = torch.optim.Adam(port.parameters(), lr=1e-3)
    ref_loss = rp.train_step(port.train(), opt, batch, 1e-4)
    assert abs(eng.last_loss() - ref_loss) < 2e-6
    np.testing.assert_allclose(_np(m.embedding.weight.grad), port.embedding.weight.grad.numpy(), rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(_np(m.embedding.weight), port.embedding.weight.detach().numpy(), rtol=1e-5, atol=5e-6)


def test_c3_igcn_step_and_topk_against_port():
    """C3 inductive configuration: IGCN (template-feature layer, edge dropout, auxiliary loss) on the Amazon-book shape"""
    from oracle import ref_port as rp
    from b200rec import ops
    g, ds, m, tr = _build("c3", {"name": "IGCN", "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1.0},
                          {"name": "IGCNTrainer", "l2_reg": 0.0, "aux_reg": 0.01})
    emb0, w0 = _np(m.embedding.weight).copy(), _np(m.w).copy()
    users, items = ds.train_pairs()
    port = rp.IGCNPort(ds.n_users, ds.n_items, users, items, emb0, 3, 0.3, m.user_map, m.item_map, w0=w0)
    ids, vals = _port_topk(port, g, 1024)
    rec = _np(tr.recommend_all("test"))[:1024]
    _check_topk(rec, ids, vals)
    m.train()
    eng = tr._engine()
    eng.step()                                    # device batch, device aux batch, device dropout mask
    torch.cuda.synchronize()
    nnz = m.feat_mat.nnz
    bits = eng.keep_bits.cpu().numpy().view(np.uint32)
    keep = ((bits[np.arange(nnz) >> 5] >> (np.arange(nnz) & 31).astype(np.uint32)) & 1).astype(bool)
    # the port's F is in the same (row-major, column-ascending) edge order as the device CSR
    r, c, _ = m.feat_mat.fwd.to_coo()
    assert np.array_equal(_np(r), port.feat_rows) and np.array_equal(_np(c), port.feat_cols)
    port.forced_keep = keep
    opt = torch.optim.Adam(port.parameters(), lr=1e-3)
    ref_loss = rp.train_step(port.train(), opt, eng.batch.cpu(), 0.0, aux_batch=eng.aux_batch.cpu(), aux_reg=0.01)
    assert abs(eng.last_loss() - ref_loss) < 2e-6
    np.testing.assert_allclose(_np(m.embedding.weight.grad), port.embedding.weight.grad.numpy(), rtol=1e-4, atol=2e-9)
    np.testing.assert_allclose(_np(m.w.grad), port.w.grad.numpy(), rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(_np(m.embedding.weight), port.embedding.weight.detach().numpy(), rtol=1e-5, atol=5e-6)
    assert ops is not None


@pytest.mark.parametrize("kind", ["LightGCN", "LightGCN4", "IGCN"])
def test_one_hop_restricted_layers_change_no_bit(kind, monkeypatch):
    """B200REC_REACH=1: forward layer L-1 is computed only within one hop of the batch (ops.mark_reach), the first
    backward hop only writes those rows and the second only gathers them; B200REC_SEL=1 (the default): the running layer
    sum is kept only at the sampled rows and G is read only there.  With a batch of 64 on the C1 graph most item
    rows lie outside the mask (the restriction really skips work); loss, gradient and the weights after three device-
    sampled steps must equal the unrestricted engine's bit for bit -- and the port's to the usual bars (LightGCN)."""
    from oracle import ref_port as rp
    L = 4 if kind == "LightGCN4" else 3
    mcfg = {"name": "IGCN", "embedding_size": 64, "n_layers": L, "dropout": 0.3, "feature_ratio": 1.0} if kind == "IGCN" \
        else {"name": "LightGCN", "embedding_size": 64, "n_layers": L}
    tcfg = {"name": "IGCNTrainer", "l2_reg": 0.0, "aux_reg": 0.01, "batch_size": 64} if kind == "IGCN" \
        else {"name": "BPRTrainer", "l2_reg": 1e-4, "batch_size": 64}
    out = {}
    for reach in ("0", "1"):
        monkeypatch.setenv("B200REC_REACH", reach)
        monkeypatch.setenv("B200REC_SEL", reach)  # and the out / addend traffic limited to the sampled rows (spmm_f32_sel)
        g, ds, m, tr = _build("c1", mcfg, tcfg)
        emb0 = _np(m.embedding.weight).copy()
        m.train()
        eng = tr._engine()
        assert (eng.reach is not None) == (reach == "1")
        losses = []
        for _ in range(3):
            eng.step()
            losses.append(eng.last_loss())
        if reach == "1":
            share = float(eng.reach[ds.n_users:].float().mean())
            assert 0.0 < share < 0.5, share                      # the mask is sparse on the item side, all-ones on the users
            assert bool(eng.reach[:ds.n_users].all())
            b = eng.batch
            assert bool(eng.reach[b[:, 1] + ds.n_users].all()) and bool(eng.reach[b[:, 2] + ds.n_users].all())
        out[reach] = (losses, m.embedding.weight.grad.clone(), m.embedding.weight.detach().clone(), eng.batch.clone())
        if reach == "1" and kind == "LightGCN":
            # the restricted engine against the CPU port on the first batch (a fresh engine: same seed, same first draw)
            g2, ds2, m2, tr2 = _build("c1", mcfg, tcfg)
            m2.train()
            e2 = tr2._engine()
            e2.step()
            users, items = ds2.train_pairs()
            port = rp.LightGCNPort(ds2.n_users, ds2.n_items, users, items, emb0, L)
            opt = torch.optim.Adam(port.parameters(), lr=1e-3)
            ref_loss = rp.train_step(port.train(), opt, e2.batch.cpu(), 1e-4)
            assert abs(e2.last_loss() - ref_loss) < 2e-6
            np.testing.assert_allclose(_np(m2.embedding.weight.grad), port.embedding.weight.grad.numpy(), rtol=1e-4, atol=1e-9)
            np.testing.assert_allclose(_np(m2.embedding.weight), port.embedding.weight.detach().numpy(), rtol=1e-5, atol=5e-6)
    assert out["0"][0] == out["1"][0]
    assert torch.equal(out["0"][3], out["1"][3])
    assert torch.equal(out["0"][1], out["1"][1])
    assert torch.equal(out["0"][2], out["1"][2])


# ------------------------------------------------------------------------------------------- C5
def test_c5_sweep_slice_tensor_core_equals_exact():
    """BASELINE config 5: 2M users x 1M items, D=128, top-20 with the real train CSR (100M entries) masked -- the tcgen05
    candidate path returns the ids and scores of the exact fp32 path on a 4 096-user slice (7 813 item tiles per user tile)"""
    import dataset as D
    from b200rec import ops, synth
    g = synth.generate_named("c4", seed=0, device=DEV, heldout=False)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": DEV, "graph": g})
    gen = torch.Generator(device=DEV).manual_seed(5)
    n = ds.n_users + ds.n_items
    rep = 0.1 * torch.randn((n, 128), device=DEV, generator=gen)
    rep[ds.n_users:, 0] += 0.02 * torch.sqrt(torch.bincount(g.train_items, minlength=ds.n_items).float())  # popularity direction
    rep[:ds.n_users, 0] += 1.0
    excl = ds.csr("train", device=DEV)
    users = torch.arange(1_000_000, 1_000_000 + 4096, dtype=torch.int64, device=DEV)
    ids1, sc1 = ops.score_topk(rep, users, rep[ds.n_users:], 20, excl_a=excl, precision=1)
    n_over = ops.score_topk.last_overflow
    ids0, sc0 = ops.score_topk(rep, users, rep[ds.n_users:], 20, excl_a=excl, precision=0)
    assert torch.equal(ids0, ids1) and torch.equal(sc0, sc1)
    assert n_over <= 41, "more than 1 % of the rows fell back to the exact path"
    # masked: no recommended item is a train item of its user; sorted by (score desc, id asc)
    ptr, idx = excl
    for j in (0, 17, 4095):
        u = int(users[j])
        row = set(idx[int(ptr[u]):int(ptr[u + 1])].tolist())
        assert not row.intersection(ids1[j].tolist())
    s = sc1.double()
    assert bool((s[:, 1:] <= s[:, :-1]).all())
