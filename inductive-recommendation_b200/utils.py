"""Drop-in for the reference's utils.py (run plumbing + graph helpers the hot path uses).

Same public names and meaning as /root/reference/utils.py: set_seed (:13-21), init_run (:24-30),
get_sparse_tensor (:33-39), generate_daj_mat (:42-50), graph_rank_nodes (:186-215), AverageMeter (:280-289),
Unbuffered (:292-305).  The DOSE augmentation builders (:71-141, :217-277) are out of scope.
"""
import os
import random
import sys

import numpy as np
import scipy.sparse as sp
import torch


def set_seed(seed=0):
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True


class Unbuffered(object):
    """file wrapper that flushes after every write (log redirection of init_run)"""

    def __init__(self, stream):
        self.stream = stream

    def write(self, data):
        self.stream.write(data)
        self.stream.flush()

    def writelines(self, datas):
        self.stream.writelines(datas)
        self.stream.flush()

    def __getattr__(self, attr):
        return getattr(self.stream, attr)


def init_run(log_path, seed):
    set_seed(seed)
    os.makedirs(log_path, exist_ok=True)
    f = Unbuffered(open(os.path.join(log_path, 'lo00gg.txt'), 'w'))
    sys.stderr = f
    sys.stdout = f


class AverageMeter:
    def __init__(self):
        self.avg, self.sum, self.count = 0., 0., 0.

    def update(self, val, n=1):
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def _train_pairs(dataset):
    if hasattr(dataset, 'train_pairs'):
        return dataset.train_pairs()
    arr = np.asarray(dataset.train_array, dtype=np.int64).reshape(-1, 2)
    return arr[:, 0], arr[:, 1]


def generate_daj_mat(dataset):
    """(U+I)^2 symmetric bipartite adjacency as scipy CSR, fp32, duplicate pairs summed."""
    users, items = _train_pairs(dataset)
    n = dataset.n_users + dataset.n_items
    r = np.concatenate([users, items + dataset.n_users])
    c = np.concatenate([items + dataset.n_users, users])
    return sp.coo_matrix((np.ones(r.shape[0]), (r, c)), shape=(n, n), dtype=np.float32).tocsr()


def get_sparse_tensor(mat, device):
    """scipy matrix -> coalesced torch sparse COO (int64 indices, fp32 values) on `device`."""
    coo = mat.tocoo()
    idx = torch.tensor(np.stack([coo.row, coo.col]), dtype=torch.int64, device=device)
    val = torch.tensor(coo.data, dtype=torch.float32, device=device)
    return torch.sparse_coo_tensor(idx, val, torch.Size(coo.shape)).coalesce()


def graph_rank_nodes(dataset, ranking_metric):
    """Rank users / items for template selection when feature_ratio < 1 (setup time, host side).
    'degree': train degree; 'sort' (and 'greedy'): column mass of the l1-row-normalised adjacency;
    'page_rank': networkx pagerank.  Returns (ranked_users, ranked_items), best first."""
    adj = generate_daj_mat(dataset)
    nu = dataset.n_users
    if ranking_metric == 'degree':
        um = np.array(np.sum(adj[:nu, :], axis=1)).squeeze()
        im = np.array(np.sum(adj[nu:, :], axis=1)).squeeze()
    elif ranking_metric in ('greedy', 'sort'):
        from sklearn.preprocessing import normalize
        nadj = normalize(adj, axis=1, norm='l1')
        um = np.array(np.sum(nadj[:, :nu], axis=0)).squeeze()
        im = np.array(np.sum(nadj[:, nu:], axis=0)).squeeze()
    elif ranking_metric == 'page_rank':
        import networkx as nx
        g = nx.Graph()
        g.add_edges_from(np.array(np.nonzero(adj)).T)
        pr = nx.pagerank(g)
        pr = np.array([pr[i] for i in range(dataset.n_users + dataset.n_items)])
        um, im = pr[:nu], pr[nu:]
    else:
        return None
    return np.argsort(um)[::-1].copy(), np.argsort(im)[::-1].copy()
