"""GPU parity tests for the contrastive models: SGL / HALF of this repo against fixtures produced by the
unmodified reference (model.py:130-365, trainer.py:432-486; oracle/make_golden.py run_contrastive_case), and the fused
InfoNCE kernel against the restated `info_nce` package (oracle/stubs/info_nce.py) in torch fp32 / fp64."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO, golden_model

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOPKS = [1, 5, 10, 15, 20]


def _np(t):
    return t.detach().cpu().numpy()


def _ref_info_nce():
    sys.path.insert(0, os.path.join(REPO, "oracle", "stubs"))
    try:
        import info_nce
    finally:
        sys.path.pop(0)
    return info_nce.info_nce


@pytest.mark.parametrize("n,d", [(37, 16), (256, 64), (2048, 64), (500, 128), (96, 8), (130, 256)])
def test_infonce_kernel_dense(n, d):
    from b200rec import ops
    ref = _ref_info_nce()
    gen = torch.Generator(device=DEV).manual_seed(n * 1000 + d)
    q = (torch.randn((n, d), generator=gen, device=DEV) * 0.1).requires_grad_(True)
    k = (q.detach() * 0.5 + 0.1 * torch.randn((n, d), generator=gen, device=DEV)).requires_grad_(True)
    loss = ops.infonce(q, k, 0.1)
    loss.backward()
    q64, k64 = q.detach().double().requires_grad_(True), k.detach().double().requires_grad_(True)
    l64 = ref(q64, k64, k64, temperature=0.1)
    l64.backward()
    assert abs(float(loss) - float(l64)) < 2e-5 * max(1.0, abs(float(l64)))
    # fp32 kernel against the fp64 formula: error budget of one fp32 pass, relative to the gradient's scale
    for got, want in ((q.grad, q64.grad), (k.grad, k64.grad)):
        scale = float(want.abs().max())
        assert float((got.double() - want).abs().max()) < 2e-5 * scale
    # and the torch fp32 evaluation of the same formula is no closer than a few ulps of that budget
    q32, k32 = q.detach().clone().requires_grad_(True), k.detach().clone().requires_grad_(True)
    l32 = ref(q32, k32, k32, temperature=0.1)
    assert abs(float(loss) - float(l32)) < 2e-5 * max(1.0, abs(float(l32)))


def test_infonce_kernel_table_form_with_duplicate_rows():
    """rows = the user column of a [B,3] batch (stride 3); a user that occurs twice receives both gradients"""
    from b200rec import ops
    ref = _ref_info_nce()
    gen = torch.Generator(device=DEV).manual_seed(5)
    n_rows, d, b = 300, 64, 128
    tq = torch.randn((n_rows, d), generator=gen, device=DEV) * 0.1
    tk = tq * 0.7 + 0.05 * torch.randn((n_rows, d), generator=gen, device=DEV)
    batch = torch.randint(0, n_rows, (b, 3), generator=gen, device=DEV, dtype=torch.int64)
    batch[5, 0] = batch[9, 0] = batch[77, 0]      # duplicates
    users = batch[:, 0]
    loss = torch.zeros(1, device=DEV)
    gq, gk = torch.zeros_like(tq), torch.zeros_like(tk)
    ops.infonce_fwd_bwd(tq, tk, loss, gq, gk, ops.infonce_workspace(b, d, DEV), temperature=0.1, loss_scale=0.25,
                        rows=batch, row_stride=3, n=b)
    tq64, tk64 = tq.double().requires_grad_(True), tk.double().requires_grad_(True)
    l64 = 0.25 * ref(tq64[users], tk64[users], tk64[users], temperature=0.1)
    l64.backward()
    assert abs(float(loss) - float(l64)) < 1e-5
    for got, want in ((gq, tq64.grad), (gk, tk64.grad)):
        assert float((got.double() - want).abs().max()) < 2e-5 * float(want.abs().max())
    untouched = torch.ones(n_rows, dtype=torch.bool, device=DEV)
    untouched[users] = False
    assert float(gq[untouched].abs().max()) == 0.0


def test_infonce_degenerate_rows():
    """an all-zero row normalises to zero (F.normalize clamps the norm at 1e-12): finite loss and gradients"""
    from b200rec import ops
    ref = _ref_info_nce()
    q = torch.randn((40, 32), device=DEV) * 0.1
    k = torch.randn((40, 32), device=DEV) * 0.1
    q[3] = 0
    k[7] = 0
    qa, ka = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
    loss = ops.infonce(qa, ka, 0.1)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(qa.grad).all() and torch.isfinite(ka.grad).all()
    assert abs(float(loss) - float(ref(q, k, k, temperature=0.1))) < 1e-5


def _assert_adam_close(actual, desired, grad_ref, lr, atol):
    """Adam's early updates are lr * m / (sqrt(v) + 1e-8): where |g| is within a few orders of eps the quotient amplifies the
    last-bit differences of g (summation order), so those entries are only held to a fraction of the step size lr."""
    solid = np.abs(grad_ref) > 1e-6
    np.testing.assert_allclose(actual[solid], desired[solid], rtol=1e-5, atol=atol)
    np.testing.assert_allclose(actual[~solid], desired[~solid], rtol=0, atol=0.25 * lr)
    assert solid.mean() > 0.8


def _keep_index(g, ds, key):
    """positions, in the dataset's train-pair order, of the edges the reference's random.sample kept for a view"""
    n_users, n_items = int(g["n_users"]), int(g["n_items"])
    idx = g[key + "_idx"]
    sel = idx[0] < n_users
    kept = idx[0][sel].astype(np.int64) * n_items + (idx[1][sel].astype(np.int64) - n_users)
    users, items = ds.train_pairs()
    allk = np.asarray(users, dtype=np.int64) * n_items + np.asarray(items, dtype=np.int64)
    order = np.argsort(allk, kind="stable")
    pos = order[np.searchsorted(allk[order], kept)]
    assert np.array_equal(allk[pos], kept)
    return np.sort(pos)


def _model(golden, name):
    g = golden(name)
    ds, m = golden_model(g, name)
    m.norm_aug_adj1 = m.generate_drop_graph(ds, keep_index=_keep_index(g, ds, "aug1"))
    if name == "sgl_tiny":
        m.norm_aug_adj2 = m.generate_drop_graph(ds, keep_index=_keep_index(g, ds, "aug2"))
    return g, ds, m


def _trainer(g, name, ds, m, **extra):
    import trainer as T
    cfg = {"name": "SGLTrainer" if name == "sgl_tiny" else "HALFTrainer", "optimizer": "Adam", "lr": float(g["lr"]),
           "l2_reg": float(g["l2_reg"]), "contrastive_reg": float(g["contrastive_reg"]), "device": DEV, "n_epochs": 1,
           "batch_size": int(g["batch"].shape[0]), "dataloader_num_workers": 0, "test_batch_size": 128, "topks": TOPKS}
    cfg.update(extra)
    return T.get_trainer(cfg, ds, m)


@pytest.mark.parametrize("name", ["sgl_tiny", "half_tiny"])
def test_views_match_reference(golden, name):
    g, ds, m = _model(golden, name)
    assert m.temperature == float(g["temperature"]) and m.aug_rate == float(g["aug_rate"])
    n_edges = len(ds.train_pairs()[0])
    for key, op in (("adj", m.norm_adj), ("aug1", m.norm_aug_adj1)) + ((("aug2", m.norm_aug_adj2),) if name == "sgl_tiny" else ()):
        r, c, v = op.to_coo()
        assert np.array_equal(np.stack([_np(r), _np(c)]), g[key + "_idx"])                 # structure: bit-exact
        np.testing.assert_allclose(_np(v), g[key + "_val"], rtol=4e-7)                     # numpy float32 pow vs correctly rounded
        if key != "adj":
            assert op.nnz == 2 * int(n_edges * m.aug_rate)                                 # utils.py:93-94
    m.eval()
    with torch.no_grad():
        np.testing.assert_allclose(_np(m.get_rep()), g["rep_eval"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(_np(m.get_aug_rep(m.norm_aug_adj1)), g["aug1_rep"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["sgl_tiny", "half_tiny"])
def test_autograd_path_two_steps(golden, name):
    g, ds, m = _model(golden, name)
    tr = _trainer(g, name, ds, m, fused=False)
    m.train()
    batch = torch.from_numpy(g["batch"]).to(DEV)
    out = m.bpr_forward(batch[:, 0].contiguous(), batch[:, 1].contiguous(), batch[:, 2].contiguous())
    assert len(out) == 5
    np.testing.assert_allclose(_np(out[0]), g["users_r"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(_np(out[3]), g["l2_norm_sq"], rtol=1e-5)
    assert abs(float(out[4]) - float(g["contrastive_loss"])) < 1e-5
    loss = tr._loss(batch, None)
    assert abs(float(loss) - float(g["loss"])) < 2e-6
    tr.opt.zero_grad()
    loss.backward()
    np.testing.assert_allclose(_np(m.embedding.weight.grad), g["grad_emb"], rtol=1e-4, atol=2e-9)
    tr.opt.step()
    _assert_adam_close(_np(m.embedding.weight), g["emb1"], g["grad_emb"], float(g["lr"]), 5e-6)
    loss2 = tr._autograd_step(batch, None)
    assert abs(loss2 - float(g["loss2"])) < 2e-6
    _assert_adam_close(_np(m.embedding.weight), g["emb2"], g["grad_emb"], float(g["lr"]), 1e-5)


@pytest.mark.parametrize("name", ["sgl_tiny", "half_tiny"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_engine_two_steps(golden, name, use_graph):
    g, ds, m = _model(golden, name)
    tr = _trainer(g, name, ds, m)
    m.train()
    eng = tr._engine()
    eng.use_graph = use_graph
    hb = torch.from_numpy(g["batch"]).pin_memory()
    eng.step(host_batch=hb)
    assert abs(eng.last_loss() - float(g["loss"])) < 2e-6
    np.testing.assert_allclose(_np(m.embedding.weight.grad), g["grad_emb"], rtol=1e-4, atol=2e-9)
    _assert_adam_close(_np(m.embedding.weight), g["emb1"], g["grad_emb"], float(g["lr"]), 5e-6)
    eng.step(host_batch=hb)
    assert abs(eng.last_loss() - float(g["loss2"])) < 2e-6
    _assert_adam_close(_np(m.embedding.weight), g["emb2"], g["grad_emb"], float(g["lr"]), 1e-5)
    eng.sync_optimizer_state()
    assert int(tr.opt.state[m.embedding.weight]["step"]) == 2


@pytest.mark.parametrize("name", ["sgl_tiny", "half_tiny"])
def test_epoch_redraws_views_and_learns(golden, name):
    g = golden(name)
    ds, m = golden_model(g, name)
    tr = _trainer(g, name, ds, m)
    m.train()
    old1 = m.norm_aug_adj1
    first = tr.train_one_epoch()
    assert m.norm_aug_adj1 is not old1 and m.aug_version == 1
    assert m.norm_aug_adj1.nnz == int(g["aug1_nnz_after_update"])                        # same size contract as the reference
    r, c, v = m.norm_aug_adj1.to_coo()
    fwd = set(zip(_np(r).tolist(), _np(c).tolist()))
    assert all((b, a) in fwd for a, b in list(fwd)[:2000])                                # symmetric
    losses = [first] + [tr.train_one_epoch() for _ in range(4)]                           # each epoch re-captures the step
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    _, metrics, _ = tr.eval("val")
    assert 0.0 <= metrics["Recall"][20] <= 1.0
