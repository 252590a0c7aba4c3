"""global top-m cosine pairs (the DOSE mining operator, model.py:503-560) against the dense torch computation"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dense(u, v, m, negate):
    un = torch.nn.functional.normalize(u.double(), dim=1)
    vn = torch.nn.functional.normalize(v.double(), dim=1) * (-1 if negate else 1)
    cos = un @ vn.T
    vals, idx = torch.topk(cos.reshape(-1), m)
    return vals, idx // v.shape[0], idx % v.shape[0], cos


@pytest.mark.parametrize("negate", [False, True])
@pytest.mark.parametrize("precision,d", [(0, 32), (1, 64)])
def test_pair_topk_global_matches_dense(negate, precision, d):
    from b200rec import mining
    gen = torch.Generator(device=DEV).manual_seed(11 + d)
    u = torch.randn((300, d), device=DEV, generator=gen)
    v = torch.randn((500, d), device=DEV, generator=gen)
    hot = torch.randn(d, device=DEV, generator=gen)
    u[7] = hot * (-1 if negate else 1)              # one user aligned with 200 items: more winners than the per-row pass keeps
    v[:200] = hot[None, :] + 0.05 * torch.randn((200, d), device=DEV, generator=gen)
    u[11] = 0                                       # a zero row has cosine 0 with everything
    m = 1500
    uu, ii, ss = mining.pair_topk_global(u, v, m, negate_items=negate, precision=precision)
    vals, ru, ri, cos = _dense(u, v, m, negate)
    assert uu.shape == ii.shape == ss.shape == (m,)
    assert float((ss.double() - vals).abs().max()) < 2e-6                     # same m values, best first
    assert int((uu == 7).sum()) >= 150                                         # the dense re-score path was needed
    got = cos[uu, ii]
    assert float((got - ss.double()).abs().max()) < 2e-6                       # every returned pair carries its own cosine
    clear = vals - vals[-1] > 1e-5                                             # winners not tied with the m-th value
    want = set(zip(ru[clear].tolist(), ri[clear].tolist()))
    assert want <= set(zip(uu.tolist(), ii.tolist()))
    key = (-ss.double()).cpu().numpy()
    assert (np.diff(key) >= 0).all()


def test_pair_topk_global_small_catalogue_and_clamp():
    from b200rec import mining
    u = torch.randn((5, 16), device=DEV)
    v = torch.randn((7, 16), device=DEV)
    uu, ii, ss = mining.pair_topk_global(u, v, 1000)                          # m larger than U*I: every pair, sorted
    assert uu.numel() == 35 and len(set(zip(uu.tolist(), ii.tolist()))) == 35
    vals, _, _, _ = _dense(u, v, 35, False)
    assert float((ss.double() - vals).abs().max()) < 2e-6
