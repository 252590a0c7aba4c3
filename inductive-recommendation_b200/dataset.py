"""Drop-in for the reference's dataset.py: the user-item graph source and the BPR sampler semantics.

Same surface as /root/reference/dataset.py -- get_dataset (:10-14), BasicDataset (:47-137, attributes n_users,
n_items, train_data, val_data, test_data, train_array, __len__, __getitem__), ProcessedDataset (:140-164),
AuxiliaryDataset (:258-273).  The raw Gowalla/Yelp/Amazon dump parsers (:167-255) are offline preprocessing and out
of scope.

B200-side additions: every dataset also keeps its interaction lists as sorted CSR arrays (`csr(split)`), which is what
the device sampler, the masking in full-rank evaluation and the graph builders read; `SyntheticDataset` wraps a
b200rec.synth graph without a text-file round trip (the 100 M-edge shapes never become Python lists).
"""
import os
import random
import sys

import numpy as np
import torch
from torch.utils.data import Dataset


def get_dataset(config):
    config = config.copy()
    cls = getattr(sys.modules[__name__], config['name'])
    return cls(config)


def _lists_to_csr(lists):
    lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=len(lists))
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    idx = np.concatenate([np.asarray(x, dtype=np.int64) for x in lists if len(x)]) if ptr[-1] else np.zeros(0, np.int64)
    return ptr, idx


class _TrackedList(list):
    """Outer list of a split (`train_data` / `val_data` / `test_data`) that counts its in-place edits, so the cached
    device CSR of the split is dropped after `dataset.test_data[u] = []` (the reference's inductive_eval,
    trainer.py:219-227) at O(1) per lookup instead of a pass over all rows.  Behaves like the list it wraps."""
    version = 0  # class-level default: unpickling (DataLoader workers) refills the list before any __init__ runs


def _counting(name):
    base = getattr(list, name)

    def method(self, *args, **kwargs):
        self.version += 1
        return base(self, *args, **kwargs)
    method.__name__ = name
    return method


for _name in ('__setitem__', '__delitem__', '__iadd__', '__imul__', 'append', 'extend', 'insert', 'pop', 'remove', 'clear',
              'reverse', 'sort'):
    setattr(_TrackedList, _name, _counting(_name))
_SPLIT_ATTRS = ('train_data', 'val_data', 'test_data')


class BasicDataset(Dataset):
    def __setattr__(self, name, value):
        if name in _SPLIT_ATTRS and type(value) is list:
            value = _TrackedList(value)  # (a shallow copy: the row lists are shared)
        object.__setattr__(self, name, value)

    def __init__(self, dataset_config):
        self.config = dataset_config
        self.name = dataset_config['name']
        self.min_interactions = dataset_config.get('min_inter')
        self.split_ratio = dataset_config.get('split_ratio')
        self.device = dataset_config['device']
        self.negative_sample_ratio = dataset_config.get('neg_ratio', 1)
        self.shuffle = dataset_config.get('shuffle', False)
        self.n_users = 0
        self.n_items = 0
        self.user_inter_lists = None
        self.train_data = None
        self.val_data = None
        self.test_data = None
        self.train_array = None
        self._csr_cache = {}

    # ---- reference sampler semantics (host side; the trainers use the device sampler by default) ----
    def __len__(self):
        return len(self.train_array)

    def __getitem__(self, index):
        """One BPR draw; `index` is ignored like in the reference: user uniform over users with a non-empty train
        list, positive uniform in that list, negatives uniform over items and rejected while in the list."""
        while True:
            user = random.randint(0, self.n_users - 1)
            if self.train_data[user]:
                break
        pos = np.random.choice(self.train_data[user])
        seen = self.train_data[user]
        out = np.empty((self.negative_sample_ratio, 3), dtype=np.int64)
        for k in range(self.negative_sample_ratio):
            neg = random.randint(0, self.n_items - 1)
            while neg in seen:
                neg = random.randint(0, self.n_items - 1)
            out[k] = (user, pos, neg)
        return out

    # ---- CSR views ----
    def csr(self, split='train', device=None, sort=True):
        """(ptr int32 [n_users+1], idx int32) of `<split>_data`, items ascending per user, cached per device."""
        lists = getattr(self, split + '_data')
        key = (split, str(device))
        hit = self._csr_cache.get(key)
        # valid while it is the same outer list and that list has not been edited: the reference's inductive_eval idiom
        # `dataset.test_data[u] = []` (trainer.py:219-227) mutates the outer list in place, which identity alone would
        # miss -- the split attributes are _TrackedList (edit counter, O(1)); a foreign list type falls back to a pass
        # over its rows.  (Editing the items INSIDE a row list in place is not detected: call invalidate_csr(split).)
        if hit is not None and hit[0] is lists and hit[2] == self._rows_fingerprint(lists):
            return hit[1]
        ptr, idx = _lists_to_csr(lists)
        if sort and idx.size:
            rows = np.repeat(np.arange(len(lists), dtype=np.int64), np.diff(ptr))
            order = np.lexsort((idx, rows))
            idx = idx[order]
        out = (torch.from_numpy(ptr.astype(np.int32)), torch.from_numpy(idx.astype(np.int32)))
        if device is not None:
            out = tuple(t.to(device) for t in out)
        self._csr_cache[key] = (lists, out, self._rows_fingerprint(lists))
        return out

    @staticmethod
    def _rows_fingerprint(lists):
        if hasattr(lists, 'ptr'):  # CSR-backed lazy lists are immutable
            return None
        if isinstance(lists, _TrackedList):
            return ('version', lists.version)
        return hash(tuple(map(id, lists))) ^ hash(tuple(map(len, lists)))

    def invalidate_csr(self, split=None):
        """drop the cached device CSR of a split (all splits if None) after its lists were edited in place"""
        for key in [k for k in self._csr_cache if split is None or k[0] == split]:
            del self._csr_cache[key]

    def train_pairs(self):
        arr = np.asarray(self.train_array, dtype=np.int64).reshape(-1, 2)
        return arr[:, 0], arr[:, 1]

    def output_dataset(self, path):
        os.makedirs(path, exist_ok=True)
        for split in ('train', 'val', 'test'):
            with open(os.path.join(path, split + '.txt'), 'w') as f:
                for user, items in enumerate(getattr(self, split + '_data')):
                    f.write(' '.join([str(user)] + [str(i) for i in items]) + '\n')


_CACHE_MAGIC = 'b200rec-lists-v1'


def _parse_lists(file_path):
    """`user item item ...` lines -> (ptr int64 [n_lines+1], idx int64 [tokens]) in file order (dataset.py:154-164 of
    the reference builds Python lists line by line; here the tokens are converted in one numpy call)."""
    with open(file_path, 'r') as f:
        lines = f.read().strip().split('\n')
    ntok = np.fromiter((len(line.split()) for line in lines), dtype=np.int64, count=len(lines))
    flat = np.array(' '.join(lines).split(), dtype=np.int64) if ntok.sum() else np.zeros(0, np.int64)
    tok_start = np.zeros(len(lines) + 1, dtype=np.int64)
    np.cumsum(ntok, out=tok_start[1:])
    keep = np.ones(flat.size, dtype=bool)
    keep[tok_start[:-1][ntok > 0]] = False                             # every non-empty line leads with its user id
    idx = flat[keep]
    ptr = np.zeros(len(lines) + 1, dtype=np.int64)
    np.cumsum(np.maximum(ntok - 1, 0), out=ptr[1:])
    return ptr, idx


def _read_lists_cached(file_path, use_cache=True):
    """Binary cache beside the text file (`<name>.txt.b200rec.npz`: ptr, idx, and the size / mtime of the text it was
    made from).  A stale or unreadable cache is ignored and rewritten; an unwritable directory just skips the cache."""
    cache = file_path + '.b200rec.npz'
    st = os.stat(file_path)
    stamp = np.array([st.st_size, st.st_mtime_ns], dtype=np.int64)
    if use_cache and os.path.exists(cache):
        try:
            with np.load(cache, allow_pickle=False) as z:
                if str(z['magic']) == _CACHE_MAGIC and np.array_equal(z['stamp'], stamp):
                    return z['ptr'], z['idx']
        except Exception:
            pass
    ptr, idx = _parse_lists(file_path)
    if use_cache:
        try:
            tmp = cache + '.tmp%d' % os.getpid()
            with open(tmp, 'wb') as f:
                np.savez(f, magic=np.array(_CACHE_MAGIC), stamp=stamp, ptr=ptr, idx=idx)
            os.replace(tmp, cache)
        except OSError:
            pass
    return ptr, idx


class ProcessedDataset(BasicDataset):
    """`<path>/{train,val,test}.txt`, one line per user: `user item item ...` (reference dataset.py:140-164), with a
    binary CSR cache written beside each file on first read (config key `cache`, default True)."""

    def __init__(self, dataset_config):
        super().__init__(dataset_config)
        path = dataset_config['path']
        self._use_cache = bool(dataset_config.get('cache', True))
        self.train_data = self.read_data(os.path.join(path, 'train.txt'))
        self.val_data = self.read_data(os.path.join(path, 'val.txt'))
        self.test_data = self.read_data(os.path.join(path, 'test.txt'))
        assert len(self.train_data) == len(self.val_data) == len(self.test_data)
        self.n_users = len(self.train_data)
        ptr, idx = _lists_to_csr(self.train_data)
        users = np.repeat(np.arange(self.n_users, dtype=np.int64), np.diff(ptr))
        self.train_array = np.stack([users, idx], axis=1).tolist()

    def read_data(self, file_path):
        ptr, idx = _read_lists_cached(file_path, self._use_cache)
        if idx.size:
            self.n_items = max(self.n_items, int(idx.max()) + 1)
        flat = idx.tolist()
        return [flat[ptr[u]:ptr[u + 1]] for u in range(len(ptr) - 1)]


class _LazyLists:
    """list-of-lists view over CSR arrays (so 100 M-edge graphs never build Python lists unless indexed)."""

    def __init__(self, ptr, idx):
        self.ptr, self.idx = ptr, idx

    def __len__(self):
        return len(self.ptr) - 1

    def __getitem__(self, u):
        return self.idx[self.ptr[u]:self.ptr[u + 1]].tolist()

    def __iter__(self):
        for u in range(len(self)):
            yield self[u]


class SyntheticDataset(BasicDataset):
    """config: {'name': 'SyntheticDataset', 'device': dev, 'graph': b200rec.synth.SynthGraph} or
    {'shape': 'c2', 'seed': 0} (BASELINE.json shapes)."""

    def __init__(self, dataset_config):
        super().__init__(dataset_config)
        from b200rec import synth
        g = dataset_config.get('graph')
        if g is None:
            g = synth.generate_named(dataset_config['shape'], seed=dataset_config.get('seed', 0),
                                     device=dataset_config.get('gen_device', 'cpu'))
        self.graph = g
        self.n_users, self.n_items = g.n_users, g.n_items
        self._np = {}
        for split in ('train', 'val', 'test'):
            ptr = getattr(g, split + '_indptr').cpu().numpy()
            idx = getattr(g, split + '_items').cpu().numpy()
            self._np[split] = (ptr, idx)
        if dataset_config.get('materialize_lists', g.train_items.numel() <= 5_000_000):
            self.train_data, self.val_data, self.test_data = (list(_LazyLists(*self._np[s])) for s in ('train', 'val', 'test'))
        else:
            self.train_data, self.val_data, self.test_data = (_LazyLists(*self._np[s]) for s in ('train', 'val', 'test'))
        self._orig = {'train': self.train_data, 'val': self.val_data, 'test': self.test_data}
        self._orig_fp = {k: self._rows_fingerprint(v) for k, v in self._orig.items()}
        self._np_csr_cache = {}
        self.train_array = None  # use train_pairs()

    def __len__(self):
        return int(self._np['train'][0][-1])

    def train_pairs(self):
        ptr, idx = self._np['train']
        return np.repeat(np.arange(self.n_users, dtype=np.int64), np.diff(ptr)), idx.astype(np.int64)

    def csr(self, split='train', device=None, sort=True):
        lists = getattr(self, split + '_data')
        # inductive_eval replaces test_data (or, written the reference's way, edits its rows in place)
        if lists is not self._orig[split] or self._rows_fingerprint(lists) != self._orig_fp[split]:
            return super().csr(split, device, sort)
        key = (split, str(device), 'np')
        hit = self._np_csr_cache.get(key)
        if hit is None:
            ptr, idx = self._np[split]  # generator output is already sorted per user
            hit = (torch.from_numpy(ptr.astype(np.int32)), torch.from_numpy(idx.astype(np.int32)))
            if device is not None:
                hit = tuple(t.to(device) for t in hit)
            self._np_csr_cache[key] = hit
        return hit


class AuxiliaryDataset(BasicDataset):
    """BPR sampler over the template (core) user/item id space used by IGCN's auxiliary loss."""

    def __init__(self, dataset, user_map, item_map):
        self.n_users = len(user_map)
        self.n_items = len(item_map)
        self.device = dataset.device
        self.negative_sample_ratio = 1
        self.length = len(dataset)
        self._csr_cache = {}
        users, items = dataset.train_pairs()
        um = np.full(dataset.n_users, -1, dtype=np.int64)
        im = np.full(dataset.n_items, -1, dtype=np.int64)
        um[np.fromiter(user_map.keys(), dtype=np.int64, count=len(user_map))] = \
            np.fromiter(user_map.values(), dtype=np.int64, count=len(user_map))
        im[np.fromiter(item_map.keys(), dtype=np.int64, count=len(item_map))] = \
            np.fromiter(item_map.values(), dtype=np.int64, count=len(item_map))
        keep = (um[users] >= 0) & (im[items] >= 0)
        tu, ti = um[users[keep]], im[items[keep]]
        order = np.argsort(tu, kind='stable')  # keeps the reference's per-user append order
        tu, ti = tu[order], ti[order]
        ptr = np.zeros(self.n_users + 1, dtype=np.int64)
        np.cumsum(np.bincount(tu, minlength=self.n_users), out=ptr[1:])
        self._ptr, self._idx = ptr, ti
        self.train_data = list(_LazyLists(ptr, ti)) if ti.size <= 5_000_000 else _LazyLists(ptr, ti)
        self.val_data = self.test_data = None

    def __len__(self):
        return self.length

    def train_pairs(self):
        return np.repeat(np.arange(self.n_users, dtype=np.int64), np.diff(self._ptr)), self._idx
