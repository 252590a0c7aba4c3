"""Host-side helpers of the drop-in utils.py / config.py against the reference's fixtures (CPU)."""
import numpy as np
import pytest

from conftest import golden_dataset, golden_maps


def test_graph_rank_nodes_sort_matches_reference_templates(golden):
    """IGCN with feature_ratio 0.5 keeps the first half of graph_rank_nodes(..., 'sort') as templates (model.py:4141-4150);
    the fixture holds the maps the unmodified reference built"""
    import utils
    g = golden("igcn_fr_tiny")
    ds = golden_dataset(g, device="cpu")
    ru, ri = utils.graph_rank_nodes(ds, "sort")
    um, im = golden_maps(g)
    assert ru[:len(um)].tolist() == list(um.keys()) and list(um.values()) == list(range(len(um)))
    assert ri[:len(im)].tolist() == list(im.keys()) and list(im.values()) == list(range(len(im)))
    assert len(um) == int(ds.n_users * 0.5) and len(im) == int(ds.n_items * 0.5)


def test_graph_rank_nodes_degree_and_unknown_metric(golden):
    import utils
    g = golden("lightgcn_tiny")
    ds = golden_dataset(g, device="cpu")
    ru, ri = utils.graph_rank_nodes(ds, "degree")
    udeg = np.diff(g["train_indptr"])
    ideg = np.bincount(g["train_items"], minlength=int(g["n_items"]))
    assert (np.diff(udeg[ru]) <= 0).all() and (np.diff(ideg[ri]) <= 0).all()
    assert sorted(ru.tolist()) == list(range(ds.n_users))
    assert utils.graph_rank_nodes(ds, "no_such_metric") is None


def test_adjacency_sums_duplicate_pairs_and_is_symmetric(golden):
    """utils.py:42-50: COO -> CSR adds duplicate (user, item) pairs; the matrix is the symmetric bipartite block form"""
    import utils

    class _DS:
        n_users, n_items = 3, 4
        train_array = [[0, 1], [0, 1], [2, 3], [1, 0]]

    a = utils.generate_daj_mat(_DS())
    assert a.shape == (7, 7) and a.dtype == np.float32
    assert a[0, 3 + 1] == 2.0 and a[3 + 1, 0] == 2.0 and a[2, 3 + 3] == 1.0
    assert (a != a.T).nnz == 0 and a[:3, :3].nnz == 0 and a[3:, 3:].nnz == 0


def test_average_meter_and_configs():
    import config
    import utils
    m = utils.AverageMeter()
    m.update(2.0, 3)
    m.update(4.0, 1)
    assert m.avg == pytest.approx(2.5) and m.count == 4
    for getter in (config.get_gowalla_config, config.get_yelp_config, config.get_amazon_config):
        triples = getter("cuda")
        names = [t[1]["name"] for t in triples]
        assert names[:3] == ["MF", "LightGCN", "IGCN"]          # the reference's list positions
        for ds_cfg, model_cfg, trainer_cfg in triples[:3]:
            assert ds_cfg["name"] == "ProcessedDataset" and trainer_cfg["batch_size"] == 2048 and trainer_cfg["topks"][-1] == 100
    c4 = config.get_synthetic_config("cuda", "c4")
    assert c4[1][1]["embedding_size"] == 128 and c4[1][1]["n_layers"] == 4


def test_config_lists_equal_the_reference():
    """every get_*_config list: same length, same positions, same keys and values as /root/reference/config.py
    (tests/golden/config_ref.json, written by oracle/make_golden.py --only-config)"""
    import json
    import os
    import config
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "config_ref.json")))
    assert set(ref) == {"get_gowalla_config", "get_yelp_config", "get_amazon_config", "get_alibaba_config", "get_ml_config"}
    for name, triples in ref.items():
        ours = [list(t) for t in getattr(config, name)("DEV")]
        assert ours == triples, name
    syn = config.get_synthetic_config("DEV", "c4")
    assert [m["name"] for _, m, _ in syn] == ["MF", "LightGCN", "IGCN", "IMF"] and syn[1][1]["embedding_size"] == 128


def test_reach_gain_and_auto_chunk():
    """host-side choices of the training engine / work plan: the expected share of item-side edges outside one hop of a
    batch (closed form against a Monte-Carlo draw of batches), and the hub-chunk rule (256-edge cap with the column-ordered
    plan of graphs >= 16 M entries, the nnz-scaled rule below)"""
    import torch
    from b200rec import graph
    rng = np.random.default_rng(3)
    n_users, n_items, batch = 5000, 800, 64
    p = np.arange(1, n_items + 1, dtype=np.float64) ** -0.9
    p /= p.sum()
    rows = [np.unique(rng.choice(n_items, size=int(k), p=p)) for k in rng.integers(1, 40, size=n_users)]
    deg = np.bincount(np.concatenate(rows), minlength=n_items)
    closed = graph.reach_gain(torch.from_numpy(deg), n_users, batch)
    mc = []
    for _ in range(200):
        hit = np.zeros(n_items, dtype=bool)
        for u in rng.integers(0, n_users, size=batch):
            hit[rows[u]] = True
        mc.append(deg[~hit].sum() / deg.sum())
    assert 0.0 < closed < 1.0 and abs(closed - float(np.mean(mc))) < 0.02
    assert graph.reach_gain(torch.from_numpy(deg), n_users, 100000) < 1e-3       # a huge batch reaches everything
    assert graph.reach_gain(torch.zeros(4, dtype=torch.int64), 10, 8) == 0.0    # no edges: nothing to skip
    assert graph.auto_chunk(2 * 1561406) == 256 and graph.auto_chunk(2 * 2984108) == 512   # C2, C3
    assert graph.auto_chunk(200_000_000) == 256 and graph.auto_chunk(graph.COLSORT_NNZ - 1) == 1024
