"""GPU: the tcgen05 / TMA score + top-K path (precision 1) must return exactly what the exact fp32 path (precision 0,
itself bit-exact against the C oracle) returns: same ids, same scores, same order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(d, k, n_items, n_users_all, n_batch, seed, scale=0.1, dup=False):
    from b200rec import ops
    rng = np.random.default_rng(seed)
    rep_u = (scale * rng.standard_normal((n_users_all, d))).astype(np.float32)
    rep_i = (scale * rng.standard_normal((n_items, d))).astype(np.float32)
    if dup:  # many exactly equal scores: the candidate band is wide, ties are broken by item id
        rep_i[: n_items // 2] = rep_i[0]
    users = rng.permutation(n_users_all)[:n_batch].astype(np.int64)
    lens = rng.integers(0, 40, n_users_all)
    ptr = np.zeros(n_users_all + 1, dtype=np.int32)
    np.cumsum(lens, out=ptr[1:])
    idx = np.concatenate([np.sort(rng.choice(n_items, l, replace=False)) for l in lens] + [np.zeros(0, int)]).astype(np.int32)
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    ru, ri, us = t(rep_u), t(rep_i), t(users)
    excl = (t(ptr), t(idx))
    banned = (n_items // 3, n_items // 3 + 17)
    out = {}
    for prec in (0, 1):
        ids, sc = ops.score_topk(ru, us, ri, k, excl_a=excl, banned=banned, precision=prec)
        torch.cuda.synchronize()
        out[prec] = (ids.cpu().numpy(), sc.cpu().numpy())
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    return ops.score_topk.last_overflow


@pytest.mark.parametrize("d,k,n_items,n_batch", [(64, 20, 3000, 200), (64, 100, 38048, 1000), (128, 20, 5001, 333),
                                                 (128, 128, 20000, 129), (64, 20, 255, 7), (64, 5, 257, 128)])
def test_tensor_core_path_equals_exact(d, k, n_items, n_batch):
    assert _case(d, k, n_items, 1500, n_batch, seed=d + k + n_items) == 0   # no row needed the exact fallback


def test_tensor_core_path_ties_and_overflow_fallback():
    assert _case(64, 20, 6000, 400, 300, seed=1, dup=True) > 0   # thousands of equal scores -> list overflow -> exact fallback


def test_tensor_core_path_large_magnitudes():
    assert _case(128, 20, 9000, 500, 256, seed=2, scale=3.0) == 0
