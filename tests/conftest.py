import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "inductive-recommendation_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _cuda_context():
    """create the CUDA context once, before the first test, and retry a box that is not ready yet (a freshly leased GPU
    has answered the very first call of a session with a transient error once); a CPU-only session does nothing"""
    import time
    import torch
    if torch.cuda.is_available():
        for attempt in range(3):
            try:
                torch.zeros(1, device="cuda")
                torch.cuda.synchronize()
                break
            except RuntimeError:
                if attempt == 2:
                    raise
                time.sleep(3.0)
    yield


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def lists_from_csr(indptr, items):
    return [items[indptr[u]:indptr[u + 1]].tolist() for u in range(len(indptr) - 1)]


def golden_dataset(g, device="cuda"):
    """SyntheticDataset over the graph stored in a golden fixture"""
    import torch
    import dataset
    from b200rec import synth
    t = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.int64))  # noqa: E731
    graph = synth.SynthGraph(int(g["n_users"]), int(g["n_items"]), t("train_indptr"), t("train_items"),
                             t("val_indptr"), t("val_items"), t("test_indptr"), t("test_items"))
    return dataset.get_dataset({"name": "SyntheticDataset", "device": device, "graph": graph})


def golden_maps(g):
    return (dict(zip(g["user_map_keys"].tolist(), g["user_map_vals"].tolist())),
            dict(zip(g["item_map_keys"].tolist(), g["item_map_vals"].tolist())))


MODEL_CFG = {
    "lightgcn_tiny": {"name": "LightGCN", "embedding_size": 64, "n_layers": 3},
    "lightgcn_d128": {"name": "LightGCN", "embedding_size": 128, "n_layers": 4},
    "mf_tiny": {"name": "MF", "embedding_size": 64},
    "igcn_tiny": {"name": "IGCN", "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1.0},
    "igcn_fr_tiny": {"name": "IGCN", "embedding_size": 64, "n_layers": 2, "dropout": 0.3, "feature_ratio": 0.5},
    "imf_tiny": {"name": "IMF", "embedding_size": 64, "n_layers": 0, "dropout": 0.3, "feature_ratio": 1.0},
    "sgl_tiny": {"name": "SGL", "embedding_size": 64, "n_layers": 3, "aug_rate": 0.8},
    "half_tiny": {"name": "HALF", "embedding_size": 64, "n_layers": 3, "aug_rate": 0.8},
    "dose_aug_tiny": {"name": "DOSE_aug", "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1,
                      "aug_num": 2000},
    "dose_drop3_tiny": {"name": "DOSE_drop3", "embedding_size": 64, "n_layers": 3, "dropout": 0.3, "feature_ratio": 1,
                        "aug_num": 2000},
}


def golden_model(g, name, device="cuda", **extra):
    """drop-in model carrying the reference's initial weights from the fixture"""
    import torch
    import model as M
    ds = golden_dataset(g, device)
    m = M.get_model(dict(MODEL_CFG[name], device=device, **extra), ds)
    with torch.no_grad():
        if name.startswith("mf"):
            m.user_embedding.weight.copy_(torch.from_numpy(g["user_emb0"]))
            m.item_embedding.weight.copy_(torch.from_numpy(g["item_emb0"]))
        else:
            m.embedding.weight.copy_(torch.from_numpy(g["emb0"]))
    return ds, m
