"""CPU: the plain-C oracle (oracle/oracle_c.c) against the reference-generated fixtures, and the stated Philox
streams against their distributional contract (dataset.py:119-131, model.py:4016-4021)."""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import ref_port as rp


def test_philox_known_answer():
    # Random123 known-answer test for philox4x32-10: counter = key = 0 / all ones / pi digits
    import ctypes as C
    lib = oc.lib()

    def run(ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        lib.oracle_philox4x32_10(c, C.c_uint32(key[0]), C.c_uint32(key[1]))
        return [int(x) for x in c]

    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("name", ["lightgcn_tiny", "lightgcn_d128"])
def test_spmm_csr_matches_reference_rep(golden, name):
    g = golden(name)
    users, items = rp.pairs_from_csr(g["train_indptr"], g["train_items"])
    a = rp.norm_adjacency(int(g["n_users"]), int(g["n_items"]), users, items)
    x = g["emb0"]
    acc = x.copy()
    for _ in range(int(g["n_layers"])):
        x = oc.spmm_csr(a.indptr, a.indices, a.data, x, chunk=64)  # small chunk: exercises the split-row order
        acc = acc + x
    rep = acc / (int(g["n_layers"]) + 1)
    np.testing.assert_allclose(rep, g["rep_eval"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["lightgcn_tiny", "igcn_tiny", "mf_tiny"])
def test_scores_and_topk_match_reference(golden, name):
    g = golden(name)
    nu = int(g["n_users"])
    if name.startswith("mf"):
        ru, ri = g["user_emb0"], g["item_emb0"]
    else:
        ru, ri = g["rep_eval"], g["rep_eval"][nu:]
    users = np.arange(nu, dtype=np.int64)
    sc = oc.scores(ru, users, ri)
    np.testing.assert_allclose(sc[:8], g["scores_head"], rtol=1e-5, atol=1e-6)
    k = 20
    tr = (g["train_indptr"], g["train_items"])
    va = (g["val_indptr"], g["val_items"])
    for split, a, b in (("train", None, None), ("val", tr, None), ("test", tr, va)):
        ids, vals = oc.mask_topk(sc, users, k, a, b)
        ref_ids, ref_val = g["topk_ids_" + split], g["topk_val_" + split]
        sep = np.ones_like(ref_ids, dtype=bool)
        d = np.abs(np.diff(ref_val, axis=1)) > 1e-6
        sep[:, 1:] &= d
        sep[:, :-1] &= d
        assert np.array_equal(ids[sep], ref_ids[sep])
        np.testing.assert_allclose(vals, ref_val, rtol=1e-5, atol=1e-6)


def test_sampler_contract(golden):
    g = golden("lightgcn_tiny")
    nu, ni = int(g["n_users"]), int(g["n_items"])
    ptr, idx = g["train_indptr"].astype(np.int32), g["train_items"].astype(np.int32)
    b = oc.bpr_sample(ptr, idx, nu, ni, seed=2021, step=0, batch=200000)
    u, p, n = b[:, 0], b[:, 1], b[:, 2]
    assert u.min() >= 0 and u.max() < nu and n.min() >= 0 and n.max() < ni
    deg = np.diff(ptr)
    assert (deg[u] > 0).all()
    train = set((np.repeat(np.arange(nu), deg) * ni + idx).tolist())
    assert all((int(a) * ni + int(c)) in train for a, c in zip(u[:5000], p[:5000]))       # positives are train items
    assert not any((int(a) * ni + int(c)) in train for a, c in zip(u[:5000], n[:5000]))   # negatives never are
    # users uniform over non-empty rows (the reference draws users uniformly, not edge-uniformly)
    cnt = np.bincount(u, minlength=nu)[deg > 0]
    expect = len(u) / (deg > 0).sum()
    assert abs(cnt.mean() - expect) < 1e-6 and cnt.std() < 3.5 * np.sqrt(expect)
    # reproducible, and a different step gives a different batch
    assert np.array_equal(b[:4096], oc.bpr_sample(ptr, idx, nu, ni, 2021, 0, 4096))
    assert not np.array_equal(b[:4096], oc.bpr_sample(ptr, idx, nu, ni, 2021, 1, 4096))


def test_empty_user_rows_are_skipped():
    ptr = np.array([0, 0, 3, 3, 5], dtype=np.int32)  # users 0 and 2 have no train items
    idx = np.array([1, 4, 7, 0, 2], dtype=np.int32)
    b = oc.bpr_sample(ptr, idx, 4, 9, seed=5, step=3, batch=4096)
    assert set(np.unique(b[:, 0]).tolist()) == {1, 3}


def test_dropout_mask_rate_and_tail():
    nnz, p = 100003, 0.3
    bits = oc.dropout_mask(nnz, p, seed=2021, step=7)
    keep = oc.unpack_bits(bits, nnz)
    assert abs(keep.mean() - (1 - p)) < 0.005
    assert (int(bits[-1]) >> (nnz % 32)) == 0  # bits past nnz are cleared
    assert np.array_equal(bits, oc.dropout_mask(nnz, p, 2021, 7))
    assert not np.array_equal(bits, oc.dropout_mask(nnz, p, 2021, 8))


@pytest.mark.parametrize("n,d", [(1, 8), (5, 16), (64, 32), (130, 64)])
def test_infonce_c_restatement_matches_package_formula(n, d):
    """oracle_infonce (plain C, double) against the restated info_nce package run in torch float64 with autograd:
    loss and both gradients, including a zero row (F.normalize's clamp)"""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "stubs"))
    try:
        from info_nce import info_nce
    finally:
        sys.path.pop(0)
    rng = np.random.default_rng(n * 100 + d)
    q = (0.1 * rng.standard_normal((n, d))).astype(np.float32)
    k = (0.5 * q + 0.1 * rng.standard_normal((n, d))).astype(np.float32)
    if n > 4:
        q[2] = 0
    loss, gq, gk = oc.infonce(q, k, 0.1)
    tq = torch.from_numpy(q).double().requires_grad_(True)
    tk = torch.from_numpy(k).double().requires_grad_(True)
    ref = info_nce(tq, tk, tk, temperature=0.1)
    ref.backward()
    assert abs(loss - float(ref)) < 1e-10
    live = np.ones(n, dtype=bool)
    if n > 4:
        live[2] = False        # autograd's subgradient at the clamp differs by convention; the clamp branch is checked below
    np.testing.assert_allclose(gq[live], tq.grad.numpy()[live], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(gk, tk.grad.numpy(), rtol=1e-8, atol=1e-12)
    if n > 4:
        assert np.isfinite(gq[2]).all()


def test_rank_metrics_c_restatement_matches_port():
    """oracle_rank_metrics against the restated calculate_metrics (oracle/ref_port.py, trainer.py:115-144)"""
    from oracle import ref_port as rp
    rng = np.random.default_rng(4)
    n_users, n_items, k = 400, 300, 50
    lens = rng.integers(0, 25, n_users)
    eval_data = [sorted(rng.choice(n_items, l, replace=False).tolist()) for l in lens]
    rec = np.stack([rng.permutation(n_items)[:k] for _ in range(n_users)]).astype(np.int32)
    ptr = np.zeros(n_users + 1, dtype=np.int32)
    np.cumsum(lens, out=ptr[1:])
    idx = np.concatenate([np.asarray(x, dtype=np.int32) for x in eval_data])
    topks = [1, 5, 10, 20, 50]
    got, counted = oc.rank_metrics(rec, ptr, idx, topks)
    want = rp.calculate_metrics(eval_data, rec, topks)
    assert counted == int((lens > 0).sum())
    for name in ("Precision", "Recall", "NDCG"):
        for kk in topks:
            assert abs(got[name][kk] - float(want[name][kk])) < 2e-6, (name, kk)
