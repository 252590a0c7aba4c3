"""ProcessedDataset: the text format of the reference (dataset.py:140-164) and the binary CSR cache beside it."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "inductive-recommendation_b200"))
import dataset as D  # noqa: E402


def _write(path, lists):
    os.makedirs(path, exist_ok=True)
    for sp in ("train", "val", "test"):
        with open(os.path.join(path, sp + ".txt"), "w") as f:
            for u, items in enumerate(lists[sp]):
                f.write(" ".join([str(u)] + [str(i) for i in items]) + "\n")


def _naive(path):
    """The reference's line-by-line parse, restated."""
    out = []
    with open(path) as f:
        for line in f.read().strip().split("\n"):
            out.append([int(t) for t in line.split(" ")[1:]])
    return out


@pytest.fixture
def lists():
    rng = np.random.default_rng(3)
    mk = lambda hi: [rng.choice(700, size=int(rng.integers(0, hi)), replace=False).tolist() for _ in range(150)]
    d = {"train": mk(40), "val": mk(5), "test": mk(8)}
    d["train"][0] = []          # a user with no interactions: the line is just the user id
    d["train"][149] = [699]
    return d


def test_parse_matches_reference_semantics(tmp_path, lists):
    p = str(tmp_path / "ds")
    _write(p, lists)
    ds = D.get_dataset({"name": "ProcessedDataset", "path": p, "device": "cpu"})
    for sp in ("train", "val", "test"):
        assert getattr(ds, sp + "_data") == _naive(os.path.join(p, sp + ".txt")) == lists[sp]
    assert ds.n_users == 150 and ds.n_items == 700
    assert ds.train_array == [[u, i] for u in range(150) for i in lists["train"][u]]
    assert len(ds) == len(ds.train_array)
    ptr, idx = ds.csr("train")
    assert ptr[-1] == len(ds.train_array)
    for u in (1, 77, 149):
        assert idx[ptr[u]:ptr[u + 1]].tolist() == sorted(lists["train"][u])


def test_cache_hit_stale_and_disabled(tmp_path, lists):
    p = str(tmp_path / "ds")
    _write(p, lists)
    cfg = {"name": "ProcessedDataset", "path": p, "device": "cpu"}
    D.get_dataset(cfg)
    cache = os.path.join(p, "train.txt.b200rec.npz")
    assert os.path.exists(cache)
    # a hit must not touch the text: make it unparsable but keep size and mtime
    txt = os.path.join(p, "val.txt")
    st = os.stat(txt)
    raw = open(txt, "rb").read()
    open(txt, "wb").write(b"x" * len(raw))
    os.utime(txt, ns=(st.st_atime_ns, st.st_mtime_ns))
    ds = D.get_dataset(cfg)
    assert ds.val_data == lists["val"]
    open(txt, "wb").write(raw)
    # a changed text invalidates the cache
    lists["train"][5] = [1, 2, 3]
    _write(p, lists)
    ds = D.get_dataset(cfg)
    assert ds.train_data[5] == [1, 2, 3]
    # a corrupt cache is ignored and rewritten
    open(cache, "wb").write(b"garbage")
    ds = D.get_dataset(cfg)
    assert ds.train_data == lists["train"]
    assert np.load(cache)["ptr"][-1] == sum(len(x) for x in lists["train"])
    # cache: False neither reads nor writes
    q = str(tmp_path / "nocache")
    _write(q, lists)
    D.get_dataset({"name": "ProcessedDataset", "path": q, "device": "cpu", "cache": False})
    assert not [f for f in os.listdir(q) if f.endswith(".npz")]


def test_output_dataset_round_trip(tmp_path, lists):
    p = str(tmp_path / "a")
    _write(p, lists)
    ds = D.get_dataset({"name": "ProcessedDataset", "path": p, "device": "cpu"})
    q = str(tmp_path / "b")
    ds.output_dataset(q)
    ds2 = D.get_dataset({"name": "ProcessedDataset", "path": q, "device": "cpu"})
    assert ds2.train_data == ds.train_data and ds2.test_data == ds.test_data and ds2.n_items == ds.n_items


def test_parser_property_random_files(tmp_path):
    """the vectorised parser against the reference's line-by-line parse on random files (users without items, single
    items, large ids)"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.lists(st.integers(0, 10 ** 6), max_size=12), min_size=1, max_size=40))
    def check(rows):
        path = str(tmp_path / "f.txt")
        with open(path, "w") as f:
            for u, items in enumerate(rows):
                f.write(" ".join([str(u)] + [str(i) for i in items]) + "\n")
        ptr, idx = D._parse_lists(path)
        got = [idx[ptr[u]:ptr[u + 1]].tolist() for u in range(len(ptr) - 1)]
        assert got == _naive(path) == rows

    check()


def test_csr_cache_follows_in_place_row_edits():
    """the reference's inductive_eval idiom `dataset.test_data[u] = []` (trainer.py:219-227) edits the outer list in
    place: the cached CSR must not survive it"""
    import dataset as D
    from b200rec import synth
    ds = D.get_dataset({"name": "SyntheticDataset", "device": "cpu", "graph": synth.generate(40, 50, 400, seed=2)})
    ptr0, idx0 = ds.csr("test")
    u = int(np.argmax(np.diff(ptr0.numpy())))
    n_u = int(ptr0[u + 1] - ptr0[u])
    assert n_u > 0
    keep = list(ds.test_data[u])
    ds.test_data[u] = []
    ptr1, idx1 = ds.csr("test")
    assert int(ptr1[-1]) == int(ptr0[-1]) - n_u and int(ptr1[u + 1] - ptr1[u]) == 0
    ds.test_data[u] = keep
    ptr2, idx2 = ds.csr("test")
    assert np.array_equal(ptr2.numpy(), ptr0.numpy()) and np.array_equal(idx2.numpy(), idx0.numpy())
    ds.test_data[u].append(int(idx0.max()))      # an edit INSIDE a row needs the explicit invalidation
    ds.invalidate_csr("test")
    assert int(ds.csr("test")[0][-1]) == int(ptr0[-1]) + 1


def test_split_lists_count_their_edits_and_survive_pickling():
    """the split attributes are list subclasses with an edit counter (the CSR cache check is O(1), not a pass over all
    rows -- 3 ms per evaluation on a Yelp-sized dataset); DataLoader workers pickle the dataset"""
    import pickle
    import dataset as D
    from b200rec import synth
    ds = D.get_dataset({"name": "SyntheticDataset", "device": "cpu", "graph": synth.generate(40, 50, 400, seed=2)})
    assert isinstance(ds.test_data, list) and type(ds.test_data).__name__ == "_TrackedList"
    v0 = ds.test_data.version
    ds.test_data[1] = []
    ds.test_data.append([3])
    assert ds.test_data.version == v0 + 2
    plain = ds.test_data.copy()                      # the reference keeps `test_data.copy()` and assigns it back
    assert type(plain) is list
    ds.test_data = plain
    assert type(ds.test_data).__name__ == "_TrackedList" and ds.test_data == plain
    again = pickle.loads(pickle.dumps(ds.test_data))
    assert again == plain and type(again).__name__ == "_TrackedList"
