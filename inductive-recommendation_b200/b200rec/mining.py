"""Global top-m (user, item) pairs by cosine similarity -- the mining step of the reference's DOSE_* models
(`cal_cos_sim_v2`, model.py:547-560: pairwise_cosine_similarity(users, items) flattened + torch.topk(aug_num);
`cal_cos_sim`, model.py:503-545, is the same selection on the NEGATED item vectors, computed on the CPU with sklearn and
split into two half-matrices).  SURVEY 8f-3: it is the evaluation kernels with a global instead of a per-row selection.

The U x I matrix is never formed: the fused score + top-K kernel (b200rec_score_topk, exact fp32 or the tcgen05 path)
returns the R best items of every user on L2-normalised rows, the m best of those U*R candidates are selected, and a
user whose R-th candidate still clears the m-th value (it may own more than R of the winners) is re-scored densely.
The DOSE model classes themselves are not built; this is the operator they would call.
"""
import torch

from . import ops

_ROW_K = 128  # per-row candidates of the first pass (the kernels' K limit)


def _unit_rows(x):
    n = x.norm(dim=1, keepdim=True)
    return torch.where(n > 0, x / n.clamp_min(1e-30), torch.zeros_like(x)).contiguous()  # zero rows stay zero (sklearn)


def pair_topk_global(rep_users, rep_items, m, negate_items=False, precision=0):
    """The m largest entries of cos(rep_users, +-rep_items): (users int64 [m], items int64 [m], cos float32 [m]), ordered by
    (cos descending, user ascending, item ascending).  Ties at the m-th value are broken arbitrarily, like torch.topk."""
    n_users, n_items = rep_users.shape[0], rep_items.shape[0]
    m = min(int(m), n_users * n_items)
    dev = rep_users.device
    un = _unit_rows(rep_users.detach().float())
    vn = _unit_rows(rep_items.detach().float())
    if negate_items:
        vn = -vn
    all_users = torch.arange(n_users, dtype=torch.int64, device=dev)
    r = min(_ROW_K, n_items)
    ids, sc = ops.score_topk(un, all_users, vn, r, precision=precision)           # [U, r], best first
    cand_u = all_users[:, None].expand(n_users, r).reshape(-1)
    cand_i, cand_s = ids.reshape(-1).long(), sc.reshape(-1)
    if r < n_items and m > 0:
        thr = torch.topk(cand_s, min(m, cand_s.numel())).values[-1]
        sat = torch.nonzero(sc[:, r - 1] >= thr).flatten()                       # rows that may own more than r winners
        if sat.numel():
            dense = ops.score_dense(un, sat.contiguous(), vn)                     # [b, I] exact fp32
            keep = torch.ones(n_users, dtype=torch.bool, device=dev)
            keep[sat] = False
            keep = keep[:, None].expand(n_users, r).reshape(-1)
            du = sat[:, None].expand(sat.numel(), n_items).reshape(-1)
            di = torch.arange(n_items, dtype=torch.int64, device=dev)[None, :].expand(sat.numel(), n_items).reshape(-1)
            cand_u = torch.cat([cand_u[keep], du])
            cand_i = torch.cat([cand_i[keep], di])
            cand_s = torch.cat([cand_s[keep], dense.reshape(-1)])
    top = torch.topk(cand_s, min(m, cand_s.numel())).indices
    u, i, s = cand_u[top], cand_i[top], cand_s[top]
    order = torch.argsort(u * n_items + i, stable=True)                           # then a stable sort by score: ties by (u, i)
    u, i, s = u[order], i[order], s[order]
    order = torch.argsort(s, descending=True, stable=True)
    return u[order], i[order], s[order]
