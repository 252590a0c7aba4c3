"""Python wrappers + torch.autograd.Functions over the C ABI (include/b200rec.h).

These are what the drop-in model.py calls where the reference calls dgl.ops.gspmm / advanced indexing:
  propagate(A, X0, L)        == stack([X0, A X0, ..., A^L X0]).mean(0)     (model.py:100-110)
  inductive_layer(F, E, ...) == dropout(F) @ E                             (model.py:4177-4190)
  gather_rows(table, idx)    == table[idx + offset, :]  (+ squared norms)  (model.py:116-119)
Everything runs on the current CUDA stream; nothing here falls back to torch math.
"""
import ctypes as C

import torch

from . import _abi
from ._abi import check, ptr, stream_ptr


def _lib():
    return _abi.load()


# ------------------------------------------------------------------------------------------------ raw calls
def spmm(op, x, keep_bits=None, post_scale=1.0, y=None, addend=None, out=None, out_scale=1.0, dst_flags=None,
         src_flags=None, out_rows=None, out_mode=0):
    """out_rows / out_mode: byte mask of the rows where out / addend matter (1: out written only there; 2: addend zero
    elsewhere, not read) -- include/b200rec.h b200rec_spmm_f32_sel"""
    _abi.require_cuda(x, keep_bits, y, addend, out, dst_flags, src_flags, out_rows)
    if out_rows is not None:
        assert keep_bits is None and out_mode in (1, 2) and x.dtype == torch.float32 and x.shape[0] >= op.n_cols
        check(_lib().b200rec_spmm_f32_sel(C.byref(op.struct()), ptr(x), x.shape[1], post_scale, ptr(y), ptr(addend), ptr(out),
                                          out_scale, ptr(dst_flags), ptr(src_flags), ptr(out_rows), out_mode, stream_ptr()),
              "spmm_f32_sel")
        return
    d = x.shape[1]
    assert x.dtype == torch.float32 and x.shape[0] >= op.n_cols
    assert op.partial is None or d <= op.max_d
    if keep_bits is None and dst_flags is None and src_flags is None and op.row_scale is None:
        blk = op.blocked_for(d)
        if blk is not None:  # table larger than L2, opt-in column blocking
            spmm_blocked(blk[0], x, blk[1], op.carry(d), post_scale=post_scale, y=y, addend=addend, out=out,
                         out_scale=out_scale)
            return
        hinted = op.hinted_for(d)
        if hinted is not None:  # table larger than L2: hot source rows kept in the persisting L2 set-aside
            op = hinted
    if dst_flags is None and src_flags is None:
        check(_lib().b200rec_spmm_f32(C.byref(op.struct()), ptr(x), d, ptr(keep_bits), post_scale, ptr(y), ptr(addend),
                                      ptr(out), out_scale, stream_ptr()), "spmm_f32")
    else:
        check(_lib().b200rec_spmm_f32_ex(C.byref(op.struct()), ptr(x), d, ptr(keep_bits), post_scale, ptr(y),
                                         ptr(addend), ptr(out), out_scale, ptr(dst_flags), ptr(src_flags),
                                         stream_ptr()), "spmm_f32_ex")


def spmm_blocked(bop, x, sweep, carry, post_scale=1.0, y=None, addend=None, out=None, out_scale=1.0):
    """The same product through a column-blocked plan (CsrOperand built with col_bounds): one pass per L2-sized block of
    source rows, per column slice of `sweep` floats of the [*, d] tables (row stride d); carry: [n_rows, d] scratch."""
    _abi.require_cuda(x, y, addend, out, carry)
    d = x.shape[1]
    assert d % sweep == 0 and carry.shape[1] == d and x.dtype == torch.float32
    off = lambda t, c0: C.c_void_p(t.data_ptr() + 4 * c0) if t is not None else C.c_void_p(0)  # noqa: E731
    for c0 in range(0, d, sweep):
        check(_lib().b200rec_spmm_f32_blocked(C.byref(bop.struct()), off(x, c0), sweep, d, post_scale, off(y, c0),
                                              off(addend, c0), off(out, c0), out_scale, off(carry, c0), stream_ptr()),
              "spmm_f32_blocked")


def live_items(op, row_flags, live_list, live_count):
    """compacted list of the plan items whose row is flagged (for spmm_live)"""
    _abi.require_cuda(row_flags, live_list, live_count)
    check(_lib().b200rec_live_items(C.byref(op.struct()), ptr(row_flags), ptr(live_list), ptr(live_count), stream_ptr()),
          "live_items")


def spmm_live(op, x, live_list, live_count, max_live, post_scale=1.0, y=None, addend=None, out=None, out_scale=1.0):
    _abi.require_cuda(x, y, addend, out, live_list, live_count)
    d = x.shape[1]
    assert x.dtype == torch.float32 and x.shape[0] >= op.n_cols
    check(_lib().b200rec_spmm_f32_live(C.byref(op.struct()), ptr(x), d, post_scale, ptr(y), ptr(addend), ptr(out), out_scale,
                                       ptr(live_list), ptr(live_count), int(max_live), stream_ptr()), "spmm_f32_live")


def propagate_fwd(op, x0, n_layers, bufs, mean_out, needed_rows=None):
    _abi.require_cuda(x0, mean_out, needed_rows)
    if n_layers > 0 and op.blocked_for(x0.shape[1]) is not None:  # the layer loop of b200rec_propagate_fwd over spmm()
        inv, src = 1.0 / (n_layers + 1), x0
        for k in range(n_layers):
            last = k == n_layers - 1
            y = None if last else bufs[k & 1]
            spmm(op, src, y=y, addend=x0 if k == 0 else mean_out, out=mean_out, out_scale=inv if last else 1.0,
                 dst_flags=needed_rows if last else None)
            src = y
        return
    check(_lib().b200rec_propagate_fwd(C.byref(op.struct()), ptr(x0), x0.shape[1], n_layers, ptr(bufs[0]), ptr(bufs[1]),
                                       ptr(mean_out), ptr(needed_rows), stream_ptr()), "propagate_fwd")


def propagate_bwd(op, g, n_layers, bufs, dx0, nonzero_rows=None):
    _abi.require_cuda(g, dx0, nonzero_rows)
    if n_layers > 0 and op.blocked_for(g.shape[1]) is not None:
        inv, src = 1.0 / (n_layers + 1), g
        for k in range(1, n_layers + 1):
            last = k == n_layers
            dst = dx0 if last else bufs[(k - 1) & 1]
            spmm(op, src, addend=g, out=dst, out_scale=inv if last else 1.0, src_flags=nonzero_rows if k == 1 else None)
            src = dst
        return
    check(_lib().b200rec_propagate_bwd(C.byref(op.struct()), ptr(g), g.shape[1], n_layers, ptr(bufs[0]), ptr(bufs[1]),
                                       ptr(dx0), ptr(nonzero_rows), stream_ptr()), "propagate_bwd")


def bpr_sample(user_ptr, user_items, n_users, n_items, seed, step, batch_size, out=None):
    if out is None:
        out = torch.empty((batch_size, 3), dtype=torch.int64, device=user_ptr.device)
    _abi.require_cuda(user_ptr, user_items, step, out)
    check(_lib().b200rec_bpr_sample(ptr(user_ptr), ptr(user_items), n_users, n_items, C.c_uint64(seed), ptr(step),
                                    batch_size, ptr(out), stream_ptr()), "bpr_sample")
    return out


def mark_rows(batch, item_offset, flags):
    """flags (uint8 [n_rows], zeroed by the caller) <- 1 at every row the batch touches"""
    _abi.require_cuda(batch, flags)
    check(_lib().b200rec_mark_rows(ptr(batch), batch.shape[0], item_offset, ptr(flags), stream_ptr()), "mark_rows")


def mark_reach(batch, item_offset, user_ptr, user_items, flags):
    """flags (uint8 [n_rows]; item half zeroed by the caller, user half kept all-ones) <- 1 at the sampled rows and at
    every train item of a sampled user: the rows within one hop of the batch"""
    _abi.require_cuda(batch, user_ptr, user_items, flags)
    check(_lib().b200rec_mark_reach(ptr(batch), batch.shape[0], item_offset, ptr(user_ptr), ptr(user_items), ptr(flags),
                                    stream_ptr()), "mark_reach")


def bpr_scratch(batch_size, d, device):
    n = int(_lib().b200rec_bpr_scratch_floats(batch_size, d))
    return torch.zeros(n, dtype=torch.float32, device=device)


def bpr_fwd_bwd(rep, batch, item_offset, l2_reg, reg_mode, g_rep, loss_out, scratch, w=None, g_w=None, loss_scale=1.0):
    _abi.require_cuda(rep, batch, g_rep, loss_out, scratch, w, g_w)
    assert batch.dtype == torch.int64 and batch.shape[1] == 3
    check(_lib().b200rec_bpr_fwd_bwd(ptr(rep), rep.shape[1], ptr(batch), batch.shape[0], item_offset, l2_reg, reg_mode,
                                     ptr(w), loss_scale, ptr(g_rep), ptr(g_w), ptr(loss_out), ptr(scratch),
                                     stream_ptr()), "bpr_fwd_bwd")


def bpr_fwd_bwd_sharded(rep, batch, item_offset, l2_reg, reg_mode, g_rep, loss_out, scratch, dots, phase, loss_weight=1.0,
                        w=None, g_w=None, loss_scale=1.0):
    _abi.require_cuda(rep, batch, g_rep, loss_out, scratch, dots, w, g_w)
    check(_lib().b200rec_bpr_fwd_bwd_sharded(ptr(rep), rep.shape[1], ptr(batch), batch.shape[0], item_offset, l2_reg,
                                             reg_mode, ptr(w), loss_scale, ptr(g_rep), ptr(g_w), ptr(loss_out),
                                             ptr(scratch), ptr(dots), phase, loss_weight, stream_ptr()),
          "bpr_fwd_bwd_sharded")


GROUP_CAP = 12288  # slots (3 * batch) the grouping holds in shared memory (csrc/bpr.cu)
GROUP_CONT = 0x80000000  # order bit 31: the position continues the row of the position before it


def bpr_grouping(batch_size, device):
    """buffers of bpr_group_rows for a [B,3] batch: (order int32 [3B], coef f32 [B])"""
    return (torch.zeros(3 * batch_size, dtype=torch.int32, device=device),
            torch.zeros(batch_size, dtype=torch.float32, device=device))


def bpr_group_rows(batch, item_offset, grouping):
    """rank the 3B (row, slot) pairs of a batch once: the ordered scatter sums a row's contributions in slot order"""
    _abi.require_cuda(batch, *grouping)
    check(_lib().b200rec_bpr_group_rows(ptr(batch), batch.shape[0], item_offset, ptr(grouping[0]), stream_ptr()),
          "bpr_group_rows")


def bpr_fwd_bwd_ordered(rep, batch, item_offset, l2_reg, reg_mode, g_rep, loss_out, scratch, grouping, dots=None, phase=0,
                        loss_weight=1.0, w=None, g_w=None, loss_scale=1.0, accumulate=False, emb0=None, l2_emb0=0.0):
    """bpr_fwd_bwd / bpr_fwd_bwd_sharded with the deterministic, aggregated scatter (one store per distinct row);
    emb0 / l2_emb0: LightGCN's layer-0 regulariser folded into the loss of the same launch"""
    _abi.require_cuda(rep, batch, g_rep, loss_out, scratch, dots, w, g_w, emb0)
    order, coef = grouping
    check(_lib().b200rec_bpr_fwd_bwd_ordered(ptr(rep), rep.shape[1], ptr(batch), batch.shape[0], item_offset, l2_reg,
                                             reg_mode, ptr(w), loss_scale, ptr(g_rep), ptr(g_w), ptr(loss_out),
                                             ptr(scratch), ptr(dots), phase, loss_weight, ptr(coef), ptr(order),
                                             1 if accumulate else 0, ptr(emb0), l2_emb0, stream_ptr()),
          "bpr_fwd_bwd_ordered")


def bpr_l2_emb0_ordered(emb0, batch, item_offset, l2_reg, g_emb0, loss_out, scratch, grouping, clear_table=None):
    """gradient of the layer-0 regulariser, one lane group per distinct row (+ its loss term unless loss_out is None: then
    bpr_fwd_bwd_ordered(emb0=...) has added it); clear_table: rows of that table zeroed in the same launch"""
    _abi.require_cuda(emb0, batch, g_emb0, loss_out, scratch, clear_table)
    check(_lib().b200rec_bpr_l2_emb0_ordered(ptr(emb0), emb0.shape[1], ptr(batch), batch.shape[0], item_offset, l2_reg,
                                             ptr(g_emb0), ptr(loss_out), ptr(scratch), ptr(grouping[0]), ptr(clear_table),
                                             stream_ptr()), "bpr_l2_emb0_ordered")


def clear_rows(batch, item_offset, table):
    """zero the rows of `table` the batch touches"""
    _abi.require_cuda(batch, table)
    check(_lib().b200rec_clear_rows(ptr(batch), batch.shape[0], item_offset, ptr(table), table.shape[1], stream_ptr()),
          "clear_rows")


def bpr_l2_emb0(emb0, batch, item_offset, l2_reg, g_emb0, loss_out, scratch):
    _abi.require_cuda(emb0, batch, g_emb0, loss_out, scratch)
    check(_lib().b200rec_bpr_l2_emb0(ptr(emb0), emb0.shape[1], ptr(batch), batch.shape[0], item_offset, l2_reg,
                                     ptr(g_emb0), ptr(loss_out), ptr(scratch), stream_ptr()), "bpr_l2_emb0")


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    _abi.require_cuda(param, grad, exp_avg, exp_avg_sq, step)
    check(_lib().b200rec_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, beta1, beta2,
                                   eps, ptr(step), stream_ptr()), "adam_step")


def step_advance(step, step_b=None, loss=None, loss_accum=None, n_batch=0):
    check(_lib().b200rec_step_advance(ptr(step), ptr(step_b), ptr(loss), ptr(loss_accum), n_batch, stream_ptr()),
          "step_advance")


def dropout_bits(nnz, p, seed, step, out=None):
    """device-drawn edge-dropout keep mask (uint32 words viewed as int32), step: device int64 scalar"""
    if out is None:
        out = torch.empty((nnz + 31) // 32, dtype=torch.int32, device=step.device)
    _abi.require_cuda(step, out)
    check(_lib().b200rec_dropout_mask(nnz, p, C.c_uint64(seed), ptr(step), ptr(out), stream_ptr()), "dropout_mask")
    return out


def infonce_workspace(n, d, device):
    return torch.empty(int(_lib().b200rec_infonce_workspace_floats(n, d)), dtype=torch.float32, device=device)


def infonce_fwd_bwd(q, k, loss_out, gq, gk, workspace, temperature=0.1, loss_scale=1.0, rows=None, row_stride=1, n=None):
    """InfoNCE (unpaired, negatives == positive keys) forward + backward; see include/b200rec.h.  Dense form: q, k [n,d],
    gq / gk overwritten.  Table form (rows = int64 index tensor, e.g. a [B,3] batch with row_stride 3): gradients are added
    into the tables gq / gk at the sampled rows."""
    _abi.require_cuda(q, k, loss_out, gq, gk, workspace, rows)
    d = q.shape[1]
    if rows is None:
        n = q.shape[0]
        assert k.shape == q.shape == gq.shape == gk.shape
    assert q.dtype == k.dtype == gq.dtype == gk.dtype == torch.float32
    assert workspace.numel() >= int(_lib().b200rec_infonce_workspace_floats(n, d))
    check(_lib().b200rec_infonce_fwd_bwd(ptr(q), ptr(k), ptr(rows), row_stride, n, d, temperature, loss_scale, ptr(loss_out),
                                         ptr(gq), ptr(gk), ptr(workspace), stream_ptr()), "infonce_fwd_bwd")


def score_dense(rep_users, users, rep_items):
    _abi.require_cuda(rep_users, users, rep_items)
    out = torch.empty((users.numel(), rep_items.shape[0]), dtype=torch.float32, device=rep_users.device)
    check(_lib().b200rec_score_dense_f32(ptr(rep_users), ptr(users), users.numel(), ptr(rep_items), rep_items.shape[0],
                                         rep_items.shape[1], ptr(out), stream_ptr()), "score_dense_f32")
    return out


def score_topk(rep_users, users, rep_items, k, excl_a=None, excl_b=None, banned=None, precision=0):
    """excl_a / excl_b: (ptr int32 [n_all_users+1], idx int32) sorted CSR by ABSOLUTE user id; banned: (lo, hi)."""
    _abi.require_cuda(rep_users, users, rep_items)
    nb, ni, d = users.numel(), rep_items.shape[0], rep_items.shape[1]
    ws_bytes = int(_lib().b200rec_score_topk_workspace(nb, ni, d, k, precision))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=rep_users.device)
    ids = torch.empty((nb, k), dtype=torch.int32, device=rep_users.device)
    sc = torch.empty((nb, k), dtype=torch.float32, device=rep_users.device)
    # an exclusion CSR without entries (e.g. an empty validation split) is the same as none
    ea = excl_a if excl_a is not None and excl_a[1].numel() else (None, None)
    eb = excl_b if excl_b is not None and excl_b[1].numel() else (None, None)
    lo, hi = banned if banned is not None else (0, 0)
    if precision == 1 and d not in (64, 128):
        precision = 0  # the tensor-core path is built for D = 64 / 128
    ovf = torch.zeros(nb, dtype=torch.int32, device=rep_users.device) if precision == 1 else None
    check(_lib().b200rec_score_topk(ptr(rep_users), ptr(users), nb, ptr(rep_items), ni, d, ptr(ea[0]), ptr(ea[1]),
                                    ptr(eb[0]), ptr(eb[1]), lo, hi, k, precision, ptr(ids), ptr(sc), ptr(ovf), ptr(ws),
                                    stream_ptr()), "score_topk")
    if precision == 1:
        bad = torch.nonzero(ovf).flatten()
        score_topk.last_overflow = int(bad.numel())  # rows that fell back to the exact path (diagnostics / tests)
        if bad.numel():  # candidate list overflow (pathological score distribution): exact path for those users
            ids2, sc2 = score_topk(rep_users, users[bad].contiguous(), rep_items, k, excl_a, excl_b, banned, precision=0)
            ids[bad], sc[bad] = ids2, sc2
    return ids, sc


def topk_global(vals, m):
    """(positions int64 [m], values [m]) of the m largest entries of a 1-D fp32 tensor, largest first, ties by position"""
    vals = vals.contiguous()
    _abi.require_cuda(vals)
    assert vals.dtype == torch.float32 and vals.dim() == 1 and 0 < m <= vals.numel()
    idx = torch.empty(m, dtype=torch.int32, device=vals.device)
    out = torch.empty(m, dtype=torch.float32, device=vals.device)
    with torch.cuda.device(vals.device):
        check(_lib().b200rec_topk_global(ptr(vals), vals.numel(), m, ptr(idx), ptr(out), stream_ptr()), "topk_global")
    return idx.long(), out


def rank_metrics(rec_ids, user0, eval_ptr, eval_idx, topks):
    """(precision, recall, ndcg) sums per cut-off and the number of users with eval items: float64 [3*len(topks)+1] on
    device.  topks ascending, each <= rec_ids.shape[1]."""
    _abi.require_cuda(rec_ids, eval_ptr, eval_idx)
    n, k = rec_ids.shape
    tk = torch.tensor(list(topks), dtype=torch.int32, device=rec_ids.device)
    blocks = (n + 127) // 128
    part = torch.empty((blocks, 3 * len(topks) + 1), dtype=torch.float64, device=rec_ids.device)
    check(_lib().b200rec_rank_metrics(ptr(rec_ids), n, k, user0, ptr(eval_ptr), ptr(eval_idx), ptr(tk), len(topks),
                                      ptr(part), stream_ptr()), "rank_metrics")
    return part.sum(0)


def hit_matrix(rec_ids, user0, eval_ptr, eval_idx):
    _abi.require_cuda(rec_ids, eval_ptr, eval_idx)
    hit = torch.empty(rec_ids.shape, dtype=torch.float32, device=rec_ids.device)
    check(_lib().b200rec_hit_matrix(ptr(rec_ids), rec_ids.shape[0], rec_ids.shape[1], user0, ptr(eval_ptr), ptr(eval_idx),
                                    ptr(hit), stream_ptr()), "hit_matrix")
    return hit


def pack_keep_bits(keep_bool):
    """bool [nnz] (CSR edge order) -> uint32 words, bit e&31 of word e>>5"""
    nnz = keep_bool.numel()
    pad = (-nnz) % 32
    k = torch.cat([keep_bool.to(torch.int64), torch.zeros(pad, dtype=torch.int64, device=keep_bool.device)])
    w = (k.view(-1, 32) << torch.arange(32, device=k.device)).sum(1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)  # two's-complement view of the 32-bit word
    return w.to(torch.int32)


# ------------------------------------------------------------------------------------------------ autograd
class _Propagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, op, n_layers):
        x0 = x0.contiguous()
        n, d = x0.shape
        bufs = [torch.empty_like(x0) if n_layers >= 2 + i else None for i in range(2)]
        out = torch.empty_like(x0)
        propagate_fwd(op, x0, n_layers, bufs, out)
        ctx.op, ctx.n_layers = op, n_layers
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n_layers = ctx.n_layers
        bufs = [torch.empty_like(g) if n_layers >= 2 + i else None for i in range(2)]
        dx0 = torch.empty_like(g)
        propagate_bwd(ctx.op, g, n_layers, bufs, dx0)
        return dx0, None, None


def propagate(op, x0, n_layers):
    """mean over layers of A^k X0, k = 0..L (differentiable w.r.t. x0)"""
    return _Propagate.apply(x0, op, n_layers)


class _InductiveLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, feat, row_scale, keep_bits, inv_keep):
        emb = emb.contiguous()
        out = torch.empty((feat.n_rows, emb.shape[1]), dtype=torch.float32, device=emb.device)
        spmm(feat.fwd.with_scales(row_scale=row_scale), emb, keep_bits=keep_bits, post_scale=inv_keep, y=out)
        ctx.feat, ctx.row_scale, ctx.keep_bits, ctx.inv_keep = feat, row_scale, keep_bits, inv_keep
        ctx.t = emb.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        demb = torch.empty((ctx.t, g.shape[1]), dtype=torch.float32, device=g.device)
        spmm(ctx.feat.bwd.with_scales(nbr_scale=ctx.row_scale), g, keep_bits=ctx.keep_bits, post_scale=ctx.inv_keep, y=demb)
        return demb, None, None, None, None


def inductive_layer(feat, emb, row_scale, keep_bits=None, inv_keep=1.0):
    """X0 = dropout(F) @ E_T with F[r, c] = row_scale[r] (model.py:4177-4190)"""
    return _InductiveLayer.apply(emb, feat, row_scale, keep_bits, inv_keep)


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, idx, offset):
        table = table.contiguous()
        idx = idx.contiguous()
        n, d = idx.numel(), table.shape[1]
        out = torch.empty((n, d), dtype=torch.float32, device=table.device)
        _abi.require_cuda(table, idx)
        check(_lib().b200rec_gather_rows(ptr(table), d, ptr(idx), offset, n, ptr(out), None, stream_ptr()), "gather_rows")
        ctx.save_for_backward(idx)
        ctx.offset, ctx.shape = offset, table.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        dt = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        check(_lib().b200rec_scatter_add_rows(ptr(dt), g.shape[1], ptr(idx), ctx.offset, idx.numel(), ptr(g),
                                              stream_ptr()), "scatter_add_rows")
        return dt, None, None


def gather_rows(table, idx, offset=0):
    """table[idx + offset, :] with a scatter-add backward (the index_put_(accumulate) of autograd)"""
    return _GatherRows.apply(table, idx, offset)


class _InfoNCE(torch.autograd.Function):
    """loss = InfoNCE(q, k) of the reference's cal_loss (model.py:206-214); gradients are produced by the same call."""

    @staticmethod
    def forward(ctx, q, k, temperature):
        q, k = q.contiguous(), k.contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=q.device)
        gq, gk = torch.empty_like(q), torch.empty_like(k)
        infonce_fwd_bwd(q, k, loss, gq, gk, infonce_workspace(q.shape[0], q.shape[1], q.device), temperature)
        ctx.save_for_backward(gq, gk)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        gq, gk = ctx.saved_tensors
        return gq * g, gk * g, None


def infonce(q, k, temperature=0.1):
    return _InfoNCE.apply(q, k, temperature)
