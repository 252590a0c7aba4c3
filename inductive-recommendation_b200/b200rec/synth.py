"""Seeded synthetic user-item interaction graphs at the BASELINE.json shapes.

There is no network for Gowalla / Yelp2018 / Amazon-book, so benchmarks and tests use power-law
bipartite graphs of the same shape: item popularity ~ rank^-a, user activity
~ rank^-a', exactly E unique train pairs, every user >= 1 train item, plus held-out val / test pairs
(about 10 % / 20 % of all interactions) that never collide with train.  The output is the per-user
item-list form the reference's ProcessedDataset parses (dataset.py:140-164); `write_processed`
emits the same `user item item ...` text files.

The generator is torch code so the 100 M-edge shapes can be produced on the GPU in about a second.  Its
randomness is integer-only (splitmix64 of a counter against integer CDF thresholds) and its set operations are
order-free, so a (shape, seed) pair names ONE graph whether it is built on the CPU or on a GPU: bench.py's own arm
and its CPU reference arm time the same graph.
"""
import os
from dataclasses import dataclass

import numpy as np
import torch

# name -> (n_users, n_items, n_train_edges, embedding_size, n_layers); BASELINE.json configs, SURVEY section 8
SHAPES = {
    "c1": (29858, 40981, 1027370, 64, 3),      # Gowalla-shaped
    "c2": (31668, 38048, 1561406, 64, 3),      # Yelp2018-shaped
    "c3": (52643, 91599, 2984108, 64, 3),      # Amazon-book-shaped
    "c4": (2000000, 1000000, 100000000, 128, 4),  # power-law 2M x 1M, 100M edges
    "tiny": (300, 500, 6000, 64, 3),
    "small": (3000, 4000, 90000, 64, 3),
}


@dataclass
class SynthGraph:
    n_users: int
    n_items: int
    # CSR by user, item ids ascending inside a row (int64 on `device`)
    train_indptr: torch.Tensor
    train_items: torch.Tensor
    val_indptr: torch.Tensor
    val_items: torch.Tensor
    test_indptr: torch.Tensor
    test_items: torch.Tensor

    def lists(self, which="train"):
        ptr = getattr(self, which + "_indptr").cpu().numpy()
        idx = getattr(self, which + "_items").cpu().numpy()
        return [idx[ptr[u]:ptr[u + 1]].tolist() for u in range(self.n_users)]


_MASK63 = (1 << 63) - 1


def _c64(v):
    """a 64-bit constant as the signed value torch's int64 arithmetic (wrapping, two's complement) takes"""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, s):
    """logical shift right of an int64 tensor (torch's >> is arithmetic)"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix64(x):
    """splitmix64 finaliser (Steele et al. 2014) on int64 tensors: integer-only, so identical on every device"""
    x = x + _c64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _c64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _c64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _stream(seed, stream, start, count, device):
    """`count` 62-bit uniform integers of stream (seed, stream), positions start .. start+count-1"""
    base = _c64((seed * 0xD1342543DE82EF95 + stream * 0x2545F4914F6CDD1D) & ((1 << 64) - 1))
    idx = torch.arange(start, start + count, dtype=torch.int64, device=device)
    return _lsr(_mix64(_mix64(idx + base) ^ base), 2)


def _powerlaw_cdf(n, a, seed, stream, device):
    """integer CDF thresholds (62-bit) of weights rank^-a, and a hash-ordered permutation of the ids.  The CDF is formed
    on the host in float64 (numpy cumsum: sequential, deterministic) and only compared as integers afterwards."""
    w = np.arange(1, n + 1, dtype=np.float64) ** (-a)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    cdf_int = np.minimum(np.floor(cdf * float(1 << 62)), float((1 << 62) - 1)).astype(np.int64)
    cdf_int[-1] = (1 << 62) - 1
    h = _stream(seed, stream, 0, n, "cpu")
    perm = torch.sort(h, stable=True).indices
    return torch.from_numpy(cdf_int).to(device), perm.to(device)


def _draw(cdf, perm, r):
    k = torch.searchsorted(cdf, r).clamp_(max=cdf.numel() - 1)
    return perm[k]


def _to_csr(keys, n_users, n_items):
    keys, _ = torch.sort(keys)
    u = torch.div(keys, n_items, rounding_mode="floor")
    i = keys - u * n_items
    counts = torch.bincount(u, minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=keys.device)
    torch.cumsum(counts, 0, out=indptr[1:])
    return indptr, i


def generate(n_users, n_items, n_train, seed=0, a_item=0.8, a_user=0.6, heldout=True, device="cpu"):
    """Return a SynthGraph with exactly `n_train` unique train pairs.  Every random choice is a splitmix64 hash of a
    counter compared against integer thresholds, and every set operation (unique, sort) is order-free, so the graph is a
    function of (shape, seed) alone: the CUDA build of it and the CPU build of it are the same graph."""
    device = torch.device(device)
    assert n_train >= n_users and n_train < n_users * n_items // 2
    ucdf, uperm = _powerlaw_cdf(n_users, a_user, seed, 1, device)
    icdf, iperm = _powerlaw_cdf(n_items, a_item, seed, 2, device)
    # one guaranteed interaction per user
    first = torch.arange(n_users, device=device, dtype=torch.int64) * n_items \
        + _draw(icdf, iperm, _stream(seed, 3, 0, n_users, device))
    keys = first
    target_total = n_train
    n_val = n_test = 0
    if heldout:
        n_val = n_train // 7
        n_test = 2 * n_train // 7
        target_total = n_train + n_val + n_test
    pos = 0
    while keys.numel() < target_total:
        m = int((target_total - keys.numel()) * 1.25) + 1024
        new = _draw(ucdf, uperm, _stream(seed, 4, pos, m, device)) * n_items \
            + _draw(icdf, iperm, _stream(seed, 5, pos, m, device))
        pos += m
        keys = torch.unique(torch.cat([keys, new]))
    # `first` must stay in train: the rest is ordered by a hash of the pair (a random order that needs no generator)
    is_first = torch.isin(keys, first)
    rest = keys[~is_first]
    order = torch.sort(_mix64(rest ^ _c64(seed * 0x9E3779B97F4A7C15 + 0x51ED270B)), stable=True).indices
    rest = rest[order]
    n_rest_train = n_train - first.numel()
    train = torch.cat([first, rest[:n_rest_train]])
    val = rest[n_rest_train:n_rest_train + n_val]
    test = rest[n_rest_train + n_val:n_rest_train + n_val + n_test]
    tp, ti = _to_csr(train, n_users, n_items)
    vp, vi = _to_csr(val, n_users, n_items)
    sp, si = _to_csr(test, n_users, n_items)
    # n_items in the reference is max id + 1 (dataset.py:159-160): make sure the last id is used
    if int(ti.max()) != n_items - 1 and not bool((ti == n_items - 1).any()):
        # replace the largest item of the highest-degree user with n_items-1 (keeps rows sorted/unique)
        deg = tp[1:] - tp[:-1]
        u = int(torch.argmax(deg))
        ti[tp[u + 1] - 1] = n_items - 1
    return SynthGraph(n_users, n_items, tp, ti, vp, vi, sp, si)


def fingerprint(graph):
    """order-sensitive 64-bit digest of the three CSRs (fixtures store it: a fixture built on one box must meet the same
    graph on another)"""
    acc = 0
    for name in ("train", "val", "test"):
        for t in (getattr(graph, name + "_indptr"), getattr(graph, name + "_items")):
            t = t.to(torch.int64)
            idx = torch.arange(t.numel(), dtype=torch.int64, device=t.device)
            h = int(_mix64(t * _c64(0x9E3779B97F4A7C15) + idx).sum().item()) & ((1 << 64) - 1)
            acc = (acc * 0x100000001B3 + h) & ((1 << 64) - 1)
    return acc


def hashed_embedding(graph, d, seed=2021, std=0.1, popularity=0.02):
    """Deterministic layer-0 table [n_users + n_items, d] fp32 for fixtures that cannot store one (C1: 18 MB): uniform
    values of standard deviation `std` from a splitmix64 hash of the element index, plus a popularity direction on
    column 0 (users +1, items + popularity * sqrt(train degree)) so that full-rank metrics sit at the popularity
    baseline instead of at chance.  Integer hash -> exact fp32 conversion, IEEE sqrt and one rounded multiply-add: the
    same bits on every machine.  numpy [N, d]."""
    n = graph.n_users + graph.n_items
    idx = torch.arange(n * d, dtype=torch.int64)
    h = _lsr(_mix64(idx + _c64(seed * 0x9E3779B97F4A7C15)), 40)             # 24 bits
    u = (h.to(torch.float32) - 8388608.0) * (1.0 / 8388608.0)                # exact: [-1, 1)
    emb = (u * np.float32(std * 3 ** 0.5)).view(n, d).numpy().copy()
    deg = torch.bincount(graph.train_items.cpu(), minlength=graph.n_items).to(torch.float32)
    emb[:graph.n_users, 0] += np.float32(1.0)
    emb[graph.n_users:, 0] += (np.float32(popularity) * torch.sqrt(deg)).numpy()
    return emb


def generate_named(name, seed=0, device="cpu", heldout=True):
    u, i, e, _, _ = SHAPES[name]
    return generate(u, i, e, seed=seed, heldout=heldout, device=device)


def write_processed(graph, path):
    """Write train.txt / val.txt / test.txt in the reference's processed format (dataset.py:40-44)."""
    os.makedirs(path, exist_ok=True)
    for which in ("train", "val", "test"):
        rows = graph.lists(which)
        with open(os.path.join(path, which + ".txt"), "w") as f:
            for u, items in enumerate(rows):
                f.write(" ".join([str(u)] + [str(i) for i in items]) + "\n")


def inductive_split(graph, old_user_frac=0.8, old_item_frac=0.8):
    """C3 protocol: ids >= frac*n are 'new' -- absent from the training graph, present at
    evaluation.  Returns (train_graph_over_old_nodes, n_old_users, n_old_items); the full `graph` is the
    enlarged dataset a model is re-pointed at for inductive_eval (trainer.py:212-253)."""
    n_old_u = int(graph.n_users * old_user_frac)
    n_old_i = int(graph.n_items * old_item_frac)

    def cut(indptr, items):
        ptr = indptr[: n_old_u + 1]
        it = items[: int(ptr[-1])]
        keep = it < n_old_i
        rows = torch.repeat_interleave(torch.arange(n_old_u, device=it.device), ptr[1:] - ptr[:-1])
        rows, it = rows[keep], it[keep]
        counts = torch.bincount(rows, minlength=n_old_u)
        p = torch.zeros(n_old_u + 1, dtype=torch.int64, device=it.device)
        torch.cumsum(counts, 0, out=p[1:])
        return p, it

    tp, ti = cut(graph.train_indptr, graph.train_items)
    vp, vi = cut(graph.val_indptr, graph.val_items)
    sp, si = cut(graph.test_indptr, graph.test_items)
    return SynthGraph(n_old_u, n_old_i, tp, ti, vp, vi, sp, si), n_old_u, n_old_i


def to_numpy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
