"""Global top-m (user, item) pairs by cosine similarity -- the mining step of the reference's DOSE_* models
(`cal_cos_sim_v2`, model.py:547-560: pairwise_cosine_similarity(users, items) flattened + torch.topk(aug_num);
`cal_cos_sim`, model.py:503-545, is the same selection on the NEGATED item vectors -- i.e. the LEAST similar pairs --
computed on the CPU with sklearn and taken as two half-matrices).  SURVEY 8f-3: it is the evaluation kernels with a
global instead of a per-row selection.

The U x I matrix is never formed: the fused score + top-K kernel (b200rec_score_topk, exact fp32 or the tcgen05 path)
returns the R best items of every user on L2-normalised rows, the m best of those U*R candidates are cut out by
b200rec_topk_global, and a user whose R-th candidate still clears the m-th value (it may own more than R of the winners)
is re-scored densely.  `DOSE_aug` / `DOSE_drop3` (model.py of this package) call `lowest_cosine_pairs`.
"""
import torch

from . import ops

_ROW_K = 128  # per-row candidates of the first pass (the kernels' K limit)


def _unit_rows(x):
    n = x.norm(dim=1, keepdim=True)
    return torch.where(n > 0, x / n.clamp_min(1e-30), torch.zeros_like(x)).contiguous()  # zero rows stay zero (sklearn)


def _top_of(cand_s, m):
    m = min(int(m), cand_s.numel())
    if m <= 0:
        return torch.zeros(0, dtype=torch.int64, device=cand_s.device)
    return ops.topk_global(cand_s.contiguous(), m)[0]


def _pair_topk_unit(un, vn, m, users, precision=0, banned=None):
    """the m largest entries of un[users] @ vn^T (unit rows given), optionally without the item range `banned`"""
    n_items = vn.shape[0]
    dev = un.device
    users = users.contiguous()
    nu = users.numel()
    allowed = n_items - ((banned[1] - banned[0]) if banned is not None else 0)
    m = min(int(m), nu * allowed)
    if m <= 0 or nu == 0:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, z.clone(), torch.zeros(0, dtype=torch.float32, device=dev)
    r = min(_ROW_K, allowed)
    ids, sc = ops.score_topk(un, users, vn, r, banned=banned, precision=precision)   # [nu, r], best first
    cand_u = users[:, None].expand(nu, r).reshape(-1)
    cand_i, cand_s = ids.reshape(-1).long(), sc.reshape(-1)
    if r < allowed:
        top = _top_of(cand_s, m)
        thr = cand_s[top[-1]]
        sat = torch.nonzero(sc[:, r - 1] >= thr).flatten()                       # rows that may own more than r winners
        if sat.numel():
            dense = ops.score_dense(un, users[sat].contiguous(), vn)              # [b, I] exact fp32
            if banned is not None:
                dense[:, banned[0]:banned[1]] = float("-inf")
            keep = torch.ones(nu, dtype=torch.bool, device=dev)
            keep[sat] = False
            keep = keep[:, None].expand(nu, r).reshape(-1)
            du = users[sat][:, None].expand(sat.numel(), n_items).reshape(-1)
            di = torch.arange(n_items, dtype=torch.int64, device=dev)[None, :].expand(sat.numel(), n_items).reshape(-1)
            cand_u = torch.cat([cand_u[keep], du])
            cand_i = torch.cat([cand_i[keep], di])
            cand_s = torch.cat([cand_s[keep], dense.reshape(-1)])
    # order the candidates by (user, item) first: b200rec_topk_global breaks ties by position, so the result is
    # (score descending, user ascending, item ascending)
    order = torch.argsort(cand_u * n_items + cand_i.clamp(min=0), stable=True)
    cand_u, cand_i, cand_s = cand_u[order], cand_i[order], cand_s[order]
    top = _top_of(cand_s, m)
    return cand_u[top], cand_i[top], cand_s[top]


def pair_topk_global(rep_users, rep_items, m, negate_items=False, precision=0):
    """The m largest entries of cos(rep_users, +-rep_items): (users int64 [m], items int64 [m], cos float32 [m]), ordered by
    (cos descending, user ascending, item ascending).  Ties at the m-th value are broken arbitrarily, like torch.topk."""
    un = _unit_rows(rep_users.detach().float())
    vn = _unit_rows(rep_items.detach().float())
    if negate_items:
        vn = -vn
    users = torch.arange(rep_users.shape[0], dtype=torch.int64, device=rep_users.device)
    return _pair_topk_unit(un, vn, m, users, precision)


def lowest_cosine_pairs(rep_users, rep_items, aug_num, reference_offsets=True, precision=0):
    """`cal_cos_sim` (model.py:503-545): -cosine of every (user, item), flattened row-major; the aug_num // 2 largest of the
    FIRST half of that vector and the aug_num // 2 largest of the SECOND half.  Returns int64 [2 * (aug_num // 2), 2]
    (user, item) rows, first-half pairs then second-half pairs, each best first.

    reference_offsets=True reproduces the reference's arithmetic for the second half: its positions j (relative to the
    half) are turned into pairs as divmod(j + aug_num // 2, n_items) instead of divmod(j + len // 2, n_items)
    (model.py:537-540), so those pairs name other cells of the matrix than the ones selected.  False gives the cells that
    were actually selected."""
    n_users, n_items = rep_users.shape[0], rep_items.shape[0]
    dev = rep_users.device
    un = _unit_rows(rep_users.detach().float())
    vn = -_unit_rows(rep_items.detach().float())
    k = int(aug_num) // 2
    half = (n_users * n_items) // 2
    u_mid, i_mid = divmod(half, n_items)
    parts = []
    for which in (0, 1):
        if which == 0:
            full = torch.arange(0, u_mid, dtype=torch.int64, device=dev)
            edge = (u_mid, (i_mid, n_items)) if i_mid > 0 else None            # row u_mid: items [0, i_mid) belong here
        else:
            full = torch.arange(u_mid + (1 if i_mid > 0 else 0), n_users, dtype=torch.int64, device=dev)
            edge = (u_mid, (0, i_mid)) if i_mid > 0 else None                  # row u_mid: items [i_mid, I)
        u, i, s = _pair_topk_unit(un, vn, k, full, precision)
        if edge is not None:
            eu = torch.tensor([edge[0]], dtype=torch.int64, device=dev)
            u2, i2, s2 = _pair_topk_unit(un, vn, k, eu, precision, banned=edge[1])
            u, i, s = torch.cat([u, u2]), torch.cat([i, i2]), torch.cat([s, s2])
            order = torch.argsort(u * n_items + i, stable=True)
            u, i, s = u[order], i[order], s[order]
            top = _top_of(s, k)
            u, i, s = u[top], i[top], s[top]
        if which == 1 and reference_offsets:
            j = u * n_items + i - half + k                                      # the reference adds aug_num // 2, not len // 2
            u, i = torch.div(j, n_items, rounding_mode="floor"), j % n_items
        parts.append(torch.stack([u, i], dim=1))
    return torch.cat(parts, dim=0)
