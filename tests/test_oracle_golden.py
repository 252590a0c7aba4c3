"""Pin the CPU oracle (oracle/ref_port.py) against the fixtures the UNMODIFIED reference produced
(tests/golden/*.npz, written by oracle/make_golden.py).  CPU-only."""
import numpy as np
import pytest
import torch

from conftest import lists_from_csr
from oracle import ref_port as rp

TOPKS = [1, 5, 10, 15, 20]


def _pairs(g):
    return rp.pairs_from_csr(g["train_indptr"], g["train_items"])


def _maps(g):
    return (dict(zip(g["user_map_keys"].tolist(), g["user_map_vals"].tolist())),
            dict(zip(g["item_map_keys"].tolist(), g["item_map_vals"].tolist())))


def _build(g, name):
    nu, ni = int(g["n_users"]), int(g["n_items"])
    users, items = _pairs(g)
    if name.startswith("lightgcn"):
        return rp.LightGCNPort(nu, ni, users, items, g["emb0"], int(g["n_layers"]))
    if name.startswith("mf"):
        return rp.MFPort(g["user_emb0"], g["item_emb0"])
    um, im = _maps(g)
    return rp.IGCNPort(nu, ni, users, items, g["emb0"], int(g["n_layers"]), float(g["dropout"]), um, im, g["w0"])


@pytest.mark.parametrize("name", ["lightgcn_tiny", "lightgcn_d128", "igcn_tiny"])
def test_norm_adj_bit_exact(golden, name):
    g = golden(name)
    users, items = _pairs(g)
    a = rp.norm_adjacency(int(g["n_users"]), int(g["n_items"]), users, items).tocoo()
    assert np.array_equal(np.stack([a.row, a.col]), g["adj_idx"])
    assert np.array_equal(a.data, g["adj_val"])  # same numpy calls -> bit-exact values
    # symmetric: backward may reuse the forward CSR
    at = a.T.tocsr()
    at.sort_indices()
    assert np.array_equal(at.tocoo().data, g["adj_val"])


@pytest.mark.parametrize("name", ["igcn_tiny", "igcn_fr_tiny", "imf_tiny"])
def test_feat_structure_and_values(golden, name):
    g = golden(name)
    users, items = _pairs(g)
    um, im = _maps(g)
    feat, row_sum = rp.build_feat(int(g["n_users"]), int(g["n_items"]), users, items, um, im)
    coo = feat.tocoo()
    assert np.array_equal(np.stack([coo.row, coo.col]), g["feat_idx"])
    assert np.array_equal(row_sum, g["row_sum"])
    assert np.array_equal(rp.feat_values(row_sum, coo.row, 1.0), g["feat_val_a1"])
    assert np.array_equal(rp.feat_values(row_sum, coo.row, float(g["alpha_after"])), g["feat_val_anneal"])


def test_template_ranking(golden):
    g = golden("igcn_fr_tiny")
    users, items = _pairs(g)
    um, im = rp.template_maps(int(g["n_users"]), int(g["n_items"]), users, items, 0.5)
    gum, gim = _maps(g)
    assert um == gum and im == gim


@pytest.mark.parametrize("name", ["lightgcn_tiny", "lightgcn_d128", "igcn_tiny", "igcn_fr_tiny", "imf_tiny"])
def test_get_rep_eval(golden, name):
    g = golden(name)
    m = _build(g, name).eval()
    with torch.no_grad():
        rep = m.get_rep().numpy()
    np.testing.assert_allclose(rep, g["rep_eval"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["lightgcn_tiny", "lightgcn_d128", "mf_tiny", "igcn_tiny", "igcn_fr_tiny", "imf_tiny"])
def test_train_step(golden, name):
    g = golden(name)
    m = _build(g, name).train()
    if "drop_keep" in g:
        m.forced_keep = np.unpackbits(g["drop_keep"])[: int(g["drop_nnz"])].astype(bool)
    batch = torch.from_numpy(g["batch"])
    bpr, loss, (ur, pr, nr, l2) = rp.bpr_loss(m, batch, float(g["l2_reg"]))
    np.testing.assert_allclose(ur.detach().numpy(), g["users_r"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(nr.detach().numpy(), g["neg_r"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(l2.detach().numpy(), g["l2_norm_sq"], rtol=1e-5)
    assert abs(float(bpr) - float(g["bpr_loss"])) < 1e-6
    if "aux_batch" in g:
        aux = rp.aux_loss(m, torch.from_numpy(g["aux_batch"]))
        assert abs(float(aux) - float(g["aux_loss"])) < 1e-6
        loss = loss + float(g["aux_reg"]) * aux
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    opt = torch.optim.Adam(m.parameters(), lr=float(g["lr"]))
    opt.zero_grad()
    loss.backward()
    if name.startswith("mf"):
        np.testing.assert_allclose(m.user_embedding.weight.grad.numpy(), g["grad_user"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(m.item_embedding.weight.grad.numpy(), g["grad_item"], rtol=1e-5, atol=1e-9)
    else:
        np.testing.assert_allclose(m.embedding.weight.grad.numpy(), g["grad_emb"], rtol=1e-4, atol=1e-9)
    if "grad_w" in g:
        np.testing.assert_allclose(m.w.grad.numpy(), g["grad_w"], rtol=1e-4, atol=1e-9)
    opt.step()
    if name.startswith("mf"):
        np.testing.assert_allclose(m.user_embedding.weight.detach().numpy(), g["user_emb1"], rtol=1e-5, atol=2e-6)
    else:
        np.testing.assert_allclose(m.embedding.weight.detach().numpy(), g["emb1"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["lightgcn_tiny", "mf_tiny", "igcn_tiny"])
def test_eval_topk_and_metrics(golden, name):
    g = golden(name)
    m = _build(g, name)
    tr = lists_from_csr(g["train_indptr"], g["train_items"])
    va = lists_from_csr(g["val_indptr"], g["val_items"])
    te = lists_from_csr(g["test_indptr"], g["test_items"])
    for split, ev in (("train", tr), ("val", va), ("test", te)):
        rec, metrics = rp.evaluate(m, tr, va, ev, split, TOPKS, test_batch_size=128)
        ref_ids, ref_val = g["topk_ids_" + split], g["topk_val_" + split]
        # ids agree wherever the reference's own adjacent scores are separated (tie order is
        # unspecified for torch.topk, trainer.py:169)
        gap_ok = np.ones_like(ref_ids, dtype=bool)
        d = np.abs(np.diff(ref_val, axis=1)) > 1e-6
        gap_ok[:, 1:] &= d
        gap_ok[:, :-1] &= d
        assert gap_ok.mean() > 0.99
        assert np.array_equal(rec[gap_ok], ref_ids[gap_ok])
        for mname in ("Precision", "Recall", "NDCG"):
            ours = np.array([metrics[mname][k] for k in TOPKS])
            np.testing.assert_allclose(ours, g["metric_%s_%s" % (mname, split)], rtol=1e-5, atol=1e-7)
    rec, metrics = rp.evaluate(m, tr, va, te, "test", TOPKS, 128, banned=(int(g["banned_lo"]), int(g["banned_hi"])))
    for mname in ("Precision", "Recall", "NDCG"):
        ours = np.array([metrics[mname][k] for k in TOPKS])
        np.testing.assert_allclose(ours, g["metric_%s_test_banned" % mname], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["sgl_tiny", "half_tiny"])
def test_contrastive_fixture_against_closed_form(golden, name):
    """SGL / HALF (model.py:130-365): the fixtures were produced through oracle/stubs/info_nce.py (the `info_nce` package
    is absent here).  Re-derive the recorded numbers WITHOUT that stub, in float64 numpy: views = mean_k A'^k E on the
    recorded edge subsets, loss_i = -s_ii/T + log(exp(s_ii/T) + sum_j exp(s_ij/T)) on L2-normalised rows."""
    import scipy.sparse as sp
    g = golden(name)
    n = int(g["n_users"]) + int(g["n_items"])
    L = int(g["n_layers"])

    def rep_of(key):
        a = sp.coo_matrix((g[key + "_val"].astype(np.float64), (g[key + "_idx"][0], g[key + "_idx"][1])), shape=(n, n)).tocsr()
        x = g["emb0"].astype(np.float64)
        acc = x.copy()
        for _ in range(L):
            x = a @ x
            acc += x
        return acc / (L + 1)

    users = g["batch"][:, 0]
    rep = rep_of("adj")
    np.testing.assert_allclose(rep, g["rep_eval"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(rep_of("aug1"), g["aug1_rep"], rtol=2e-5, atol=1e-7)
    if name == "sgl_tiny":
        q, k = rep_of("aug1")[users], rep_of("aug2")[users]
    else:
        q, k = rep[users], rep_of("aug1")[users]
    qn = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
    kn = k / np.maximum(np.linalg.norm(k, axis=1, keepdims=True), 1e-12)
    s = qn @ kn.T / float(g["temperature"])
    pos = np.diag(s)
    loss = np.mean(-pos + np.log(np.exp(pos) + np.exp(s).sum(1)))
    assert abs(loss - float(g["contrastive_loss"])) < 2e-5
    # the view keeps exactly int(E * aug_rate) train pairs, both directions (utils.py:91-103)
    n_edges = int(g["train_indptr"][-1])
    assert g["aug1_idx"].shape[1] == 2 * int(n_edges * float(g["aug_rate"])) == int(g["aug1_nnz_after_update"])
    # total loss recorded by the trainer arithmetic (trainer.py:447-452)
    ur = g["users_r"].astype(np.float64)
    total = float(g["bpr_loss"]) + float(g["l2_reg"]) * float(np.mean(g["l2_norm_sq"])) + float(g["contrastive_reg"]) * loss
    assert abs(total - float(g["loss"])) < 1e-5 and ur.shape == (len(users), 64)


@pytest.mark.parametrize("name", ["sgl_tiny", "half_tiny"])
def test_contrastive_port_matches_reference_steps(golden, name):
    """oracle/ref_port.py SGLPort + contrastive_train_step against the unmodified reference's two recorded steps"""
    g = golden(name)
    nu, ni = int(g["n_users"]), int(g["n_items"])
    users, items = _pairs(g)

    def view(key):
        idx = g[key + "_idx"]
        sel = idx[0] < nu
        return idx[0][sel].astype(np.int64), idx[1][sel].astype(np.int64) - nu

    views = [view("aug1")] + ([view("aug2")] if name == "sgl_tiny" else [])
    port = rp.SGLPort(nu, ni, users, items, g["emb0"], int(g["n_layers"]), views, half=(name == "half_tiny")).train()
    a = port.views[0].to_sparse_coo().coalesce()
    assert np.array_equal(a.indices().numpy(), g["aug1_idx"]) and np.array_equal(a.values().numpy(), g["aug1_val"])
    opt = torch.optim.Adam(port.parameters(), lr=float(g["lr"]))
    batch = torch.from_numpy(g["batch"])
    loss, con = rp.contrastive_train_step(port, opt, batch, float(g["l2_reg"]), float(g["contrastive_reg"]))
    assert abs(loss - float(g["loss"])) < 1e-6 and abs(con - float(g["contrastive_loss"])) < 1e-6
    np.testing.assert_allclose(port.embedding.weight.grad.numpy(), g["grad_emb"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(port.embedding.weight.detach().numpy(), g["emb1"], rtol=1e-5, atol=2e-6)
    loss2, _ = rp.contrastive_train_step(port, opt, batch, float(g["l2_reg"]), float(g["contrastive_reg"]))
    assert abs(loss2 - float(g["loss2"])) < 1e-6
    np.testing.assert_allclose(port.embedding.weight.detach().numpy(), g["emb2"], rtol=1e-5, atol=4e-6)


def test_port_matches_reference_at_c1_size():
    """the port against the BASELINE-sized fixture c1_ref (29 858 x 40 981, 1 027 370 edges, D=64, L=3, batch 2048) that
    oracle/make_golden.py --only-c1 recorded from the unmodified reference: representation, one training step, and the
    full Recall/NDCG@20 of eval('test') (trainer.py:115-210).  The inputs are regenerated: the generator and the hashed
    layer-0 table are deterministic, the fixture stores the graph's fingerprint."""
    from conftest import load_golden
    from b200rec import synth
    f = load_golden("c1_ref")
    g = synth.generate_named("c1", seed=0)
    assert synth.fingerprint(g) == int(f["fingerprint"])
    emb0 = synth.hashed_embedding(g, 64)
    assert abs(float(np.abs(emb0.astype(np.float64)).sum()) - float(f["emb0_checksum"])) < 1e-6
    ptr, idx = g.train_indptr.numpy(), g.train_items.numpy()
    users, items = rp.pairs_from_csr(ptr, idx)
    m = rp.LightGCNPort(g.n_users, g.n_items, users, items, emb0, 3).eval()
    coo = m.adj_sp.tocoo()
    assert coo.nnz == int(f["adj_nnz"])
    pos = f["adj_pos"]
    assert np.array_equal(np.stack([coo.row[pos], coo.col[pos]]), f["adj_idx_at"])
    np.testing.assert_allclose(coo.data[pos], f["adj_val_at"], rtol=3e-7)
    with torch.no_grad():
        rep = m.get_rep().numpy()
    np.testing.assert_allclose(rep[f["rows_kept"]], f["rep_eval_rows"], rtol=1e-5, atol=1e-7)
    # full-rank evaluation from the cached representation (the reference re-propagates per batch; same numbers)
    n_u = g.n_users
    rec = np.empty((n_u, 20), dtype=np.int64)
    vptr, vidx = g.val_indptr.numpy(), g.val_items.numpy()
    for s in range(0, n_u, 2048):
        e = min(n_u, s + 2048)
        sc = rep[s:e] @ rep[n_u:].T
        for u in range(s, e):
            sc[u - s, idx[ptr[u]:ptr[u + 1]]] = -np.inf
            sc[u - s, vidx[vptr[u]:vptr[u + 1]]] = -np.inf
        rec[s:e] = rp.topk_tiebreak(sc, 20)[0]
    metrics = rp.calculate_metrics(g.lists("test"), rec, TOPKS)
    for name in ("Precision", "Recall", "NDCG"):
        ours = np.array([metrics[name][k] for k in TOPKS])
        np.testing.assert_allclose(ours, f["metric_%s_test" % name], rtol=1e-5, atol=1e-7)
    ref_ids, ref_val = f["topk_ids_test"], f["topk_val_test"]
    d = np.abs(np.diff(ref_val, axis=1)) > 1e-6
    sep = np.ones_like(ref_ids, dtype=bool)
    sep[:, 1:] &= d
    sep[:, :-1] &= d
    assert sep.mean() > 0.97 and np.array_equal(rec[: ref_ids.shape[0]][sep], ref_ids[sep])
    # one training step on the reference's recorded batch
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=float(f["lr"]))
    loss = rp.train_step(m, opt, torch.from_numpy(f["batch"]), float(f["l2_reg"]))
    assert abs(loss - float(f["loss"])) < 1e-6
    np.testing.assert_allclose(m.embedding.weight.grad.numpy()[f["rows_kept"]], f["grad_rows"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(m.embedding.weight.detach().numpy()[f["rows_kept"]], f["emb1_rows"], rtol=1e-5, atol=2e-6)


def test_fast_adjacency_builder_equals_the_scipy_one():
    """bench.py's CPU arm builds the 100M-edge adjacency with norm_adjacency_from_csr: bit-identical to norm_adjacency"""
    from b200rec import synth
    for g in (synth.generate(300, 500, 6000, seed=7), synth.generate_named("c1", seed=0)):
        ptr, idx = g.train_indptr.numpy(), g.train_items.numpy()
        users, items = rp.pairs_from_csr(ptr, idx)
        a = rp.norm_adjacency(g.n_users, g.n_items, users, items)
        b = rp.norm_adjacency_from_csr(g.n_users, g.n_items, ptr, idx)
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
        assert a.data.dtype == b.data.dtype == np.float32 and np.array_equal(a.data, b.data)
