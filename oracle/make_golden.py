"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/*.py) on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (where /root/reference is mounted); the
fixtures it writes are committed so they travel to the GPU box, which has no /root/reference.

    python oracle/make_golden.py            # rewrites tests/golden/{lightgcn,igcn,igcn_fr,imf,mf}_tiny.npz
    python oracle/make_golden.py --only-c1  # the BASELINE-sized case c1_ref.npz (about 3 minutes of CPU)

The tiny fixtures carry their graph inside the file (they were drawn with the round-1 generator); c1_ref names its
graph by (shape, seed) and stores the generator's fingerprint.

The three third-party leaves the reference imports but this image lacks (dgl, info_nce,
torchmetrics) are satisfied by oracle/stubs (contract: SURVEY.md Appendix B).  Everything recorded
here is an output of the reference's own classes:
  * graph: LightGCN.generate_graph (model.py:89-98) -> norm_adj COO (row-major, coalesced)
  * get_rep (model.py:100-110 / :4188-4200), bpr_forward (:112-120 / :4046-4052)
  * BPRTrainer / IGCNTrainer loss + backward + Adam step (trainer.py:412-429 / :531-561)
  * BasicTrainer.eval top-K ids + metrics (trainer.py:146-210, :115-144)
  * IGCN feat_mat / row_sum / anneal (model.py:4127-4175), dropout mask draw (:4016-4028)
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, "inductive-recommendation_b200"))
from b200rec import synth  # noqa: E402  (graph generator only; no kernels)

sys.path.remove(os.path.join(REPO, "inductive-recommendation_b200"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, "stubs"))

import torch  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import dataset as ref_dataset  # noqa: E402
    import model as ref_model  # noqa: E402
    import trainer as ref_trainer  # noqa: E402
    import utils as ref_utils  # noqa: E402

TOPKS = [1, 5, 10, 15, 20]
CPU = torch.device("cpu")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def coo_of(sp):
    sp = sp.coalesce()
    return sp.indices().numpy().copy(), sp.values().detach().numpy().copy()


def run_case(tag, graph, model_cfg, trainer_cfg, out_dir, seed=2021, batch=256):
    tmp = tempfile.mkdtemp()
    synth.write_processed(graph, tmp)
    ds = quiet(ref_dataset.get_dataset, {"name": "ProcessedDataset", "path": tmp, "device": CPU})
    ref_utils.set_seed(seed)
    model = quiet(ref_model.get_model, dict(model_cfg, device=CPU), ds)
    tr = quiet(ref_trainer.get_trainer,
               dict(trainer_cfg, device=CPU, dataloader_num_workers=0, topks=TOPKS, n_epochs=1,
                    batch_size=batch, test_batch_size=128), ds, model)
    out = {}
    out["n_users"], out["n_items"] = ds.n_users, ds.n_items
    out["train_indptr"] = graph.train_indptr.numpy()
    out["train_items"] = graph.train_items.numpy()
    out["val_indptr"], out["val_items"] = graph.val_indptr.numpy(), graph.val_items.numpy()
    out["test_indptr"], out["test_items"] = graph.test_indptr.numpy(), graph.test_items.numpy()
    name = model_cfg["name"]
    if hasattr(model, "norm_adj"):
        out["adj_idx"], out["adj_val"] = coo_of(model.norm_adj)
    if name == "MF":
        out["user_emb0"] = model.user_embedding.weight.detach().numpy().copy()
        out["item_emb0"] = model.item_embedding.weight.detach().numpy().copy()
    else:
        out["emb0"] = model.embedding.weight.detach().numpy().copy()
    is_igcn = name in ("IGCN", "IMF")
    if is_igcn:
        out["feat_idx"], out["feat_val_a1"] = coo_of(model.feat_mat)
        out["row_sum"] = model.row_sum.numpy().copy()
        out["user_map_keys"] = np.array(list(model.user_map.keys()), dtype=np.int64)
        out["user_map_vals"] = np.array(list(model.user_map.values()), dtype=np.int64)
        out["item_map_keys"] = np.array(list(model.item_map.keys()), dtype=np.int64)
        out["item_map_vals"] = np.array(list(model.item_map.values()), dtype=np.int64)
        out["w0"] = model.w.detach().numpy().copy()

    # ---- eval-mode representation + full-rank evaluation (no dropout) ----
    model.eval()
    if name != "MF":
        with torch.no_grad():
            out["rep_eval"] = model.get_rep().numpy().copy()
    users_all = torch.arange(ds.n_users, dtype=torch.int64)
    with torch.no_grad():
        out["scores_head"] = model.predict(users_all[:8]).numpy().copy()
    for split in ("train", "val", "test"):
        _, metrics, _ = tr.eval(split)
        # the reference does not keep rec_items; recompute exactly as trainer.py:146-170 does
        rec = []
        with torch.no_grad():
            for (users,) in tr.test_user_loader:
                scores = model.predict(users)
                if split != "train":
                    ex_u, ex_i = [], []
                    for ui, u in enumerate(users.tolist()):
                        items = ds.train_data[u]
                        if split == "test":
                            items = items + ds.val_data[u]
                        ex_u.extend([ui] * len(items))
                        ex_i.extend(items)
                    scores[ex_u, ex_i] = -np.inf
                v, it = torch.topk(scores, k=max(TOPKS))
                rec.append((v.numpy(), it.numpy()))
        out["topk_ids_" + split] = np.concatenate([r[1] for r in rec], 0)
        out["topk_val_" + split] = np.concatenate([r[0] for r in rec], 0)
        for m in ("Precision", "Recall", "NDCG"):
            out["metric_%s_%s" % (m, split)] = np.array([metrics[m][k] for k in TOPKS], dtype=np.float64)
    # banned-item pass used by inductive_eval (trainer.py:236-238)
    n_old_items = int(ds.n_items * 0.8)
    _, metrics, _ = tr.eval("test", banned_items=np.arange(n_old_items, ds.n_items))
    out["banned_lo"], out["banned_hi"] = n_old_items, ds.n_items
    for m in ("Precision", "Recall", "NDCG"):
        out["metric_%s_test_banned" % m] = np.array([metrics[m][k] for k in TOPKS], dtype=np.float64)

    # ---- one training step on a recorded batch (training mode) ----
    model.train()
    ref_utils.set_seed(seed + 1)
    rows = [ds[0] for _ in range(batch)]  # BasicDataset.__getitem__ (dataset.py:119-131)
    b = torch.tensor(np.stack(rows)[:, 0, :], dtype=torch.int64)
    out["batch"] = b.numpy().copy()
    users, pos, neg = b[:, 0], b[:, 1], b[:, 2]
    if is_igcn:
        # record the dropout draw NGCF.dropout_sp_mat makes (model.py:4019-4021): CPU generator
        torch.manual_seed(seed + 2)
        nnz = model.feat_mat._nnz()
        rnd = torch.rand(nnz)
        keep = torch.floor((1 - model.dropout) + rnd).type(torch.bool)
        out["drop_keep"] = np.packbits(keep.numpy())
        out["drop_nnz"] = nnz
        out["dropout"] = model.dropout
        torch.manual_seed(seed + 2)
    ur, pr, nr, l2 = model.bpr_forward(users, pos, neg)
    out["users_r"], out["pos_r"], out["neg_r"], out["l2_norm_sq"] = (
        ur.detach().numpy().copy(), pr.detach().numpy().copy(), nr.detach().numpy().copy(), l2.detach().numpy().copy())
    F = torch.nn.functional
    bpr = F.softplus((ur * nr).sum(1) - (ur * pr).sum(1)).mean()
    l2_reg = trainer_cfg["l2_reg"]
    loss = bpr + l2_reg * l2.mean()
    if name in ("IGCN", "IMF") and trainer_cfg["name"] == "IGCNTrainer":
        aux_ds = tr.aux_dataloader.dataset
        ref_utils.set_seed(seed + 3)
        arows = [aux_ds[0] for _ in range(batch)]
        ab = torch.tensor(np.stack(arows)[:, 0, :], dtype=torch.int64)
        out["aux_batch"] = ab.numpy().copy()
        au, ap, an = ab[:, 0], ab[:, 1], ab[:, 2]
        tu = len(model.user_map)
        e_u, e_p, e_n = model.embedding(au), model.embedding(ap + tu), model.embedding(an + tu)
        ps = (e_u * e_p * model.w[None, :]).sum(1)
        ns = (e_u * e_n * model.w[None, :]).sum(1)
        aux = F.softplus(ns - ps).mean()
        out["aux_loss"] = float(aux)
        loss = loss + trainer_cfg["aux_reg"] * aux
    out["bpr_loss"] = float(bpr)
    out["loss"] = float(loss)
    tr.opt.zero_grad()
    loss.backward()
    if name == "MF":
        out["grad_user"] = model.user_embedding.weight.grad.numpy().copy()
        out["grad_item"] = model.item_embedding.weight.grad.numpy().copy()
    else:
        out["grad_emb"] = model.embedding.weight.grad.numpy().copy()
    if is_igcn:
        out["grad_w"] = model.w.grad.numpy().copy()
    tr.opt.step()
    if name == "MF":
        out["user_emb1"] = model.user_embedding.weight.detach().numpy().copy()
        out["item_emb1"] = model.item_embedding.weight.detach().numpy().copy()
    else:
        out["emb1"] = model.embedding.weight.detach().numpy().copy()
    if is_igcn:
        out["w1"] = model.w.detach().numpy().copy()
        model.feat_mat_anneal()
        out["alpha_after"] = model.alpha
        out["feat_val_anneal"] = coo_of(model.feat_mat)[1]
    out["lr"] = trainer_cfg["lr"]
    out["l2_reg"] = l2_reg
    out["aux_reg"] = trainer_cfg.get("aux_reg", 0.0)
    out["n_layers"] = model_cfg.get("n_layers", 0)
    out["topks"] = np.array(TOPKS)
    path = os.path.join(out_dir, tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def run_inductive_case(tag, full_graph, out_dir, seed=2021):
    """IGCN built on the OLD nodes only, then re-pointed at the enlarged dataset the way IGCN.load does
    (model.py:4213-4220: generate_feat(is_updating=True) + update_feat_mat) and evaluated with the reference's
    inductive_eval (trainer.py:212-253); its six printed result lines are the fixture."""
    old_graph, n_old_u, n_old_i = synth.inductive_split(full_graph, 0.8, 0.8)
    tmp_old, tmp_new = tempfile.mkdtemp(), tempfile.mkdtemp()
    synth.write_processed(old_graph, tmp_old)
    synth.write_processed(full_graph, tmp_new)
    ds_old = quiet(ref_dataset.get_dataset, {"name": "ProcessedDataset", "path": tmp_old, "device": CPU})
    ds_new = quiet(ref_dataset.get_dataset, {"name": "ProcessedDataset", "path": tmp_new, "device": CPU})
    ref_utils.set_seed(seed)
    cfg = {"name": "IGCN", "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1.0, "device": CPU}
    model = quiet(ref_model.get_model, cfg, ds_old)
    out = {"emb0": model.embedding.weight.detach().numpy().copy(), "n_old_users": ds_old.n_users,
           "n_old_items": ds_old.n_items, "old_n_items_attr": ds_old.n_items}
    for name, g in (("old", old_graph), ("new", full_graph)):
        for split in ("train", "val", "test"):
            out["%s_%s_indptr" % (name, split)] = getattr(g, split + "_indptr").numpy()
            out["%s_%s_items" % (name, split)] = getattr(g, split + "_items").numpy()
    out["new_n_users"], out["new_n_items"] = ds_new.n_users, ds_new.n_items
    # re-point the model at the enlarged dataset (what a driver has to do; the reference ships none)
    model.config["dataset"] = ds_new
    model.n_users, model.n_items = ds_new.n_users, ds_new.n_items
    model.norm_adj = model.generate_graph(ds_new)
    model.feat_mat, _, _, model.row_sum = model.generate_feat(ds_new, is_updating=True)
    model.update_feat_mat()
    model.eval()
    with torch.no_grad():
        out["rep_new"] = model.get_rep().numpy().copy()
    out["feat_idx_new"], out["feat_val_new"] = coo_of(model.feat_mat)
    tr = quiet(ref_trainer.get_trainer,
               {"name": "IGCNTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 0.0, "aux_reg": 0.01, "device": CPU,
                "dataloader_num_workers": 0, "topks": TOPKS, "n_epochs": 1, "batch_size": 256, "test_batch_size": 128},
               ds_new, model)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tr.inductive_eval(ds_old.n_users, ds_old.n_items)
    lines = [l for l in buf.getvalue().splitlines() if "result." in l]
    assert len(lines) == 6, lines
    out["inductive_lines"] = np.array(lines)
    path = os.path.join(out_dir, tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    for l in lines:
        print("   ", l)


def run_contrastive_case(tag, graph, model_name, out_dir, seed=2021, batch=256):
    """SGL (model.py:130-243, SGLTrainer trainer.py:432-459) / HALF (model.py:246-365, HALFTrainer :460-486): LightGCN
    propagation on the full graph and on edge-dropped views + InfoNCE between the views' user rows."""
    tmp = tempfile.mkdtemp()
    synth.write_processed(graph, tmp)
    ds = quiet(ref_dataset.get_dataset, {"name": "ProcessedDataset", "path": tmp, "device": CPU})
    ref_utils.set_seed(seed)
    mcfg = {"name": model_name, "embedding_size": 64, "n_layers": 3, "aug_rate": 0.8}
    tcfg = {"name": model_name + "Trainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4, "contrastive_reg": 0.1}
    model = quiet(ref_model.get_model, dict(mcfg, device=CPU), ds)
    tr = quiet(ref_trainer.get_trainer,
               dict(tcfg, device=CPU, dataloader_num_workers=0, topks=TOPKS, n_epochs=1, batch_size=batch,
                    test_batch_size=128), ds, model)
    out = {"n_users": ds.n_users, "n_items": ds.n_items, "train_indptr": graph.train_indptr.numpy(),
           "train_items": graph.train_items.numpy(), "val_indptr": graph.val_indptr.numpy(),
           "val_items": graph.val_items.numpy(), "test_indptr": graph.test_indptr.numpy(),
           "test_items": graph.test_items.numpy()}
    out["adj_idx"], out["adj_val"] = coo_of(model.norm_adj)
    out["aug1_idx"], out["aug1_val"] = coo_of(model.norm_aug_adj1)
    if model_name == "SGL":
        out["aug2_idx"], out["aug2_val"] = coo_of(model.norm_aug_adj2)
    out["emb0"] = model.embedding.weight.detach().numpy().copy()
    model.eval()
    with torch.no_grad():
        out["rep_eval"] = model.get_rep().numpy().copy()
        out["aug1_rep"] = model.get_aug_rep(model.norm_aug_adj1).numpy().copy()
    model.train()
    ref_utils.set_seed(seed + 1)
    rows = [ds[0] for _ in range(batch)]
    b = torch.tensor(np.stack(rows)[:, 0, :], dtype=torch.int64)
    out["batch"] = b.numpy().copy()
    users, pos, neg = b[:, 0], b[:, 1], b[:, 2]
    F = torch.nn.functional
    # trainer.py:441-456 restated on the recorded batch (the reference draws its batches inside the DataLoader loop)
    ur, pr, nr, l2, con = model.bpr_forward(users, pos, neg)
    bpr = F.softplus((ur * nr).sum(1) - (ur * pr).sum(1)).mean()
    loss = bpr + tcfg["l2_reg"] * l2.mean() + tcfg["contrastive_reg"] * con.mean()
    out["users_r"], out["l2_norm_sq"] = ur.detach().numpy().copy(), l2.detach().numpy().copy()
    out["bpr_loss"], out["contrastive_loss"], out["loss"] = float(bpr), float(con), float(loss)
    tr.opt.zero_grad()
    loss.backward()
    out["grad_emb"] = model.embedding.weight.grad.numpy().copy()
    tr.opt.step()
    out["emb1"] = model.embedding.weight.detach().numpy().copy()
    # a second step on the same batch pins the optimizer state hand-over
    ur, pr, nr, l2, con = model.bpr_forward(users, pos, neg)
    loss2 = (F.softplus((ur * nr).sum(1) - (ur * pr).sum(1)).mean() + tcfg["l2_reg"] * l2.mean()
             + tcfg["contrastive_reg"] * con.mean())
    tr.opt.zero_grad()
    loss2.backward()
    tr.opt.step()
    out["loss2"] = float(loss2)
    out["emb2"] = model.embedding.weight.detach().numpy().copy()
    # epoch end: both views are redrawn (model.py:231-237); only the size contract is recorded
    quiet(model.update_aug_adj)
    out["aug1_nnz_after_update"] = int(model.norm_aug_adj1._nnz())
    out["lr"], out["l2_reg"], out["contrastive_reg"] = tcfg["lr"], tcfg["l2_reg"], tcfg["contrastive_reg"]
    out["aug_rate"], out["n_layers"], out["temperature"] = mcfg["aug_rate"], mcfg["n_layers"], 0.1
    path = os.path.join(out_dir, tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def run_baseline_case(tag, shape, out_dir, n_rows_kept=2048, n_users_kept=1024, batch=2048, seed=2021):
    """A BASELINE.json-sized case (C1: 29 858 x 40 981, 1 027 370 edges, LightGCN L=3 D=64, batch 2048): the unmodified
    reference's normalised adjacency, eval-mode representation, one recorded training step and eval('val'/'test') with
    Recall/NDCG@20.  The inputs are regenerated by the tests (synth.generate_named + synth.hashed_embedding are
    deterministic; the fixture stores the graph's fingerprint), the outputs are kept at `n_rows_kept` sampled table rows
    and `n_users_kept` users so that the file stays at a few MB."""
    graph = synth.generate_named(shape, seed=0)
    tmp = tempfile.mkdtemp()
    synth.write_processed(graph, tmp)
    ds = quiet(ref_dataset.get_dataset, {"name": "ProcessedDataset", "path": tmp, "device": CPU})
    ref_utils.set_seed(seed)
    _, _, _, d, n_layers = synth.SHAPES[shape]
    model = quiet(ref_model.get_model, {"name": "LightGCN", "embedding_size": d, "n_layers": n_layers, "device": CPU}, ds)
    tcfg = {"name": "BPRTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4}
    tr = quiet(ref_trainer.get_trainer, dict(tcfg, device=CPU, dataloader_num_workers=0, topks=TOPKS, n_epochs=1,
                                             batch_size=batch, test_batch_size=512), ds, model)
    emb0 = synth.hashed_embedding(graph, d)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(emb0))     # an input, like the graph
    n = ds.n_users + ds.n_items
    out = {"shape": shape, "fingerprint": np.uint64(synth.fingerprint(graph)), "n_users": ds.n_users, "n_items": ds.n_items,
           "emb0_checksum": np.float64(np.abs(emb0.astype(np.float64)).sum())}
    idx, val = coo_of(model.norm_adj)
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(val.shape[0], size=8192, replace=False))
    out["adj_nnz"], out["adj_pos"], out["adj_idx_at"], out["adj_val_at"] = val.shape[0], pos, idx[:, pos], val[pos]
    model.eval()
    with torch.no_grad():
        rep = model.get_rep().numpy()
    # ---- one training step on a batch drawn by the reference's own sampler ----
    model.train()
    ref_utils.set_seed(seed + 1)
    rows = [ds[0] for _ in range(batch)]
    b = torch.tensor(np.stack(rows)[:, 0, :], dtype=torch.int64)
    out["batch"] = b.numpy().copy()
    touched = np.unique(np.concatenate([b[:, 0].numpy(), ds.n_users + b[:, 1].numpy(), ds.n_users + b[:, 2].numpy()]))
    keep = np.unique(np.concatenate([rng.choice(touched, size=n_rows_kept // 2, replace=False),
                                     rng.choice(n, size=n_rows_kept // 2, replace=False)]))
    out["rows_kept"] = keep
    out["rep_eval_rows"] = rep[keep].copy()
    out["rep_eval_abs_sum"] = np.float64(np.abs(rep.astype(np.float64)).sum())
    users, p_, n_ = b[:, 0], b[:, 1], b[:, 2]
    ur, pr, nr, l2 = model.bpr_forward(users, p_, n_)
    F = torch.nn.functional
    bpr = F.softplus((ur * nr).sum(1) - (ur * pr).sum(1)).mean()
    loss = bpr + tcfg["l2_reg"] * l2.mean()
    out["bpr_loss"], out["loss"] = float(bpr), float(loss)
    tr.opt.zero_grad()
    loss.backward()
    g = model.embedding.weight.grad.numpy()
    out["grad_rows"] = g[keep].copy()
    out["grad_abs_sum"] = np.float64(np.abs(g.astype(np.float64)).sum())
    tr.opt.step()
    out["emb1_rows"] = model.embedding.weight.detach().numpy()[keep].copy()
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(emb0))     # evaluate the recorded input, not the stepped one
    # ---- full-rank evaluation (trainer.py:146-210) ----
    model.eval()
    for split in ("val", "test"):
        _, metrics, _ = tr.eval(split)
        for m in ("Precision", "Recall", "NDCG"):
            out["metric_%s_%s" % (m, split)] = np.array([metrics[m][k] for k in TOPKS], dtype=np.float64)
    with torch.no_grad():
        us = torch.arange(n_users_kept, dtype=torch.int64)
        scores = model.predict(us)
        ex_u, ex_i = [], []
        for ui, u in enumerate(us.tolist()):
            items = ds.train_data[u] + ds.val_data[u]
            ex_u.extend([ui] * len(items))
            ex_i.extend(items)
        scores[ex_u, ex_i] = -np.inf
        v, it = torch.topk(scores, k=max(TOPKS))
    out["topk_ids_test"], out["topk_val_test"] = it.numpy(), v.numpy()
    out["lr"], out["l2_reg"], out["n_layers"], out["topks"] = tcfg["lr"], tcfg["l2_reg"], n_layers, np.array(TOPKS)
    path = os.path.join(out_dir, tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    print("   recall@20 test", out["metric_Recall_test"][-1], "ndcg@20", out["metric_NDCG_test"][-1], "loss", out["loss"])


def run_dose_case(tag, graph, model_name, trainer_name, out_dir, aug_num=2000, seed=2021, batch=256):
    """DOSE_aug (model.py:367-613) / DOSE_drop3 (:2544-2863) with DOSEaugTrainer / DOSEdropTrainer (trainer.py:255-353):
    IGCN + a second propagation on a graph edited by similarity mining (cal_cos_sim: the aug_num (user, item) pairs with
    the LOWEST cosine similarity are added to / removed from the train graph) + InfoNCE between the two views' user rows.
    Recorded: the mined pairs, the edited adjacency, one training step (two dropout draws: one per view) and the
    eval-mode representation."""
    tmp = tempfile.mkdtemp()
    synth.write_processed(graph, tmp)
    ds = quiet(ref_dataset.get_dataset, {"name": "ProcessedDataset", "path": tmp, "device": CPU})
    ref_utils.set_seed(seed)
    mcfg = {"name": model_name, "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1, "aug_num": aug_num}
    tcfg = {"name": trainer_name, "optimizer": "Adam", "lr": 1e-3, "l2_reg": 0., "contrastive_reg": 1e-1, "aux_reg": 0.001}
    model = quiet(ref_model.get_model, dict(mcfg, device=CPU), ds)       # mines with the initial weights (eval of __init__: training mode!)
    tr = quiet(ref_trainer.get_trainer, dict(tcfg, device=CPU, dataloader_num_workers=0, topks=TOPKS, n_epochs=1,
                                             batch_size=batch, test_batch_size=128), ds, model)
    out = {"n_users": ds.n_users, "n_items": ds.n_items, "aug_num": aug_num}
    for split in ("train", "val", "test"):
        out[split + "_indptr"] = getattr(graph, split + "_indptr").numpy()
        out[split + "_items"] = getattr(graph, split + "_items").numpy()
    out["emb0"] = model.embedding.weight.detach().numpy().copy()
    out["w0"] = model.w.detach().numpy().copy()
    out["adj_idx"], out["adj_val"] = coo_of(model.norm_adj)
    # mining on a known representation: eval mode (no dropout), so that the test can reproduce the input exactly
    model.eval()
    with torch.no_grad():
        rep = model.get_def_rep().numpy().copy()
    out["rep_eval"] = rep
    with torch.no_grad():
        pairs = np.array(quiet(model.cal_cos_sim), dtype=np.int64)          # [aug_num, 2] in the reference's order
    out["mined_pairs"] = pairs
    # cosine values the selection saw (sklearn float32) at the mined pairs and the two cut values, for tie-aware checks
    from sklearn.metrics.pairwise import cosine_similarity
    cos = cosine_similarity(rep[:ds.n_users], -rep[ds.n_users:]).astype(np.float32).reshape(-1)
    half = cos.shape[0] // 2
    k = aug_num // 2
    out["cut_first"] = np.float32(np.sort(cos[:half])[-k])
    out["cut_second"] = np.float32(np.sort(cos[half:])[-k])
    # the edited graph built from THOSE pairs (generate_aug_graph / generate_drop_graph re-run the mining: same input, same pairs)
    with torch.no_grad():
        edited = model.generate_aug_graph(ds) if model_name == "DOSE_aug" else model.generate_drop_graph(ds)
    model.norm_aug_adj = edited
    out["aug_idx"], out["aug_val"] = coo_of(edited)
    # ---- one training step on a recorded batch, with the two recorded dropout draws ----
    model.train()
    ref_utils.set_seed(seed + 1)
    rows = [ds[0] for _ in range(batch)]
    b = torch.tensor(np.stack(rows)[:, 0, :], dtype=torch.int64)
    aux_ds = tr.aux_dataloader.dataset
    ref_utils.set_seed(seed + 3)
    ab = torch.tensor(np.stack([aux_ds[0] for _ in range(batch)])[:, 0, :], dtype=torch.int64)
    out["batch"], out["aux_batch"] = b.numpy().copy(), ab.numpy().copy()
    nnz = model.feat_mat._nnz()
    torch.manual_seed(seed + 2)
    keeps = [torch.floor((1 - model.dropout) + torch.rand(nnz)).type(torch.bool) for _ in range(2)]  # def view, then aug view
    out["drop_keep_def"], out["drop_keep_aug"] = np.packbits(keeps[0].numpy()), np.packbits(keeps[1].numpy())
    out["drop_nnz"], out["dropout"] = nnz, model.dropout
    torch.manual_seed(seed + 2)
    F = torch.nn.functional
    users, pos, neg = b[:, 0], b[:, 1], b[:, 2]
    ur, pr, nr, l2, con = model.bpr_forward(users, pos, neg)
    bpr = F.softplus((ur * nr).sum(1) - (ur * pr).sum(1)).mean()
    au, ap, an = ab[:, 0], ab[:, 1], ab[:, 2]
    tu = len(model.user_map)
    e_u, e_p, e_n = model.embedding(au), model.embedding(ap + tu), model.embedding(an + tu)
    aux = F.softplus((e_u * e_n * model.w[None, :]).sum(1) - (e_u * e_p * model.w[None, :]).sum(1)).mean()
    loss = bpr + tcfg["l2_reg"] * l2.mean() + tcfg["aux_reg"] * aux + tcfg["contrastive_reg"] * con.mean()   # trainer.py:290-291
    out["users_r"] = ur.detach().numpy().copy()
    out["bpr_loss"], out["aux_loss"], out["contrastive_loss"], out["loss"] = float(bpr), float(aux), float(con), float(loss)
    tr.opt.zero_grad()
    loss.backward()
    out["grad_emb"] = model.embedding.weight.grad.numpy().copy()
    out["grad_w"] = model.w.grad.numpy().copy()
    tr.opt.step()
    out["emb1"] = model.embedding.weight.detach().numpy().copy()
    out["w1"] = model.w.detach().numpy().copy()
    out["lr"], out["l2_reg"], out["aux_reg"], out["contrastive_reg"] = tcfg["lr"], tcfg["l2_reg"], tcfg["aux_reg"], tcfg["contrastive_reg"]
    out["n_layers"] = 3
    path = os.path.join(out_dir, tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "loss", out["loss"], "con", out["contrastive_loss"],
          "edited nnz", out["aug_val"].shape[0], "vs", out["adj_val"].shape[0])


def main():
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "--only-dose" in sys.argv:
        g = synth.generate(300, 500, 6000, seed=7)
        run_dose_case("dose_aug_tiny", g, "DOSE_aug", "DOSEaugTrainer", out_dir)
        run_dose_case("dose_drop3_tiny", g, "DOSE_drop3", "DOSEdropTrainer", out_dir)
        return
    if "--only-config" in sys.argv:  # the reference's config lists (positions + hyper-parameters), device = "DEV"
        import importlib.util
        import json
        spec = importlib.util.spec_from_file_location("ref_config", "/root/reference/config.py")
        cfg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cfg)
        names = ["get_gowalla_config", "get_yelp_config", "get_amazon_config", "get_alibaba_config", "get_ml_config"]
        json.dump({n: [list(t) for t in getattr(cfg, n)("DEV")] for n in names},
                  open(os.path.join(out_dir, "config_ref.json"), "w"), indent=0)
        return
    if "--only-c1" in sys.argv:
        run_baseline_case("c1_ref", "c1", out_dir)
        return
    if "--only-contrastive" in sys.argv:
        g = synth.generate(300, 500, 6000, seed=7)
        run_contrastive_case("sgl_tiny", g, "SGL", out_dir)
        run_contrastive_case("half_tiny", g, "HALF", out_dir)
        return
    if "--only-inductive" in sys.argv:
        run_inductive_case("igcn_inductive", synth.generate(400, 600, 9000, seed=21), out_dir)
        return
    g = synth.generate(300, 500, 6000, seed=7)
    bpr = {"name": "BPRTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4}
    igcn = {"name": "IGCNTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 0.0, "aux_reg": 0.01}
    run_case("lightgcn_tiny", g, {"name": "LightGCN", "embedding_size": 64, "n_layers": 3}, bpr, out_dir)
    run_case("mf_tiny", g, {"name": "MF", "embedding_size": 64}, dict(bpr, l2_reg=1e-3), out_dir)
    run_case("igcn_tiny", g, {"name": "IGCN", "embedding_size": 64, "n_layers": 3, "dropout": 0.3,
                              "feature_ratio": 1.0}, igcn, out_dir)
    run_case("igcn_fr_tiny", g, {"name": "IGCN", "embedding_size": 64, "n_layers": 2, "dropout": 0.3,
                                 "feature_ratio": 0.5}, dict(igcn, l2_reg=1e-5), out_dir)
    run_case("imf_tiny", g, {"name": "IMF", "embedding_size": 64, "n_layers": 0, "dropout": 0.3,
                             "feature_ratio": 1.0}, igcn, out_dir)
    # a D=128 / L=4 case (the C4 kernel configuration) on a graph with hub rows
    g2 = synth.generate(400, 300, 9000, seed=11, a_item=1.1)
    run_case("lightgcn_d128", g2, {"name": "LightGCN", "embedding_size": 128, "n_layers": 4}, bpr, out_dir)
    run_inductive_case("igcn_inductive", synth.generate(400, 600, 9000, seed=21), out_dir)
    run_contrastive_case("sgl_tiny", g, "SGL", out_dir)
    run_contrastive_case("half_tiny", g, "HALF", out_dir)


if __name__ == "__main__":
    main()
