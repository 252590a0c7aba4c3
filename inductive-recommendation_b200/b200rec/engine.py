"""Fused BPR training step, captured once as a CUDA graph and replayed per batch.

One replay = one iteration of BPRTrainer / IGCNTrainer.train_one_epoch (trainer.py:412-429, :531-558):
  device sampling (or a host batch copied in) -> [edge dropout + inductive layer] -> L x SpMM + layer mean
  -> fused BPR loss/grad -> L x SpMM backward (Horner) -> [transposed inductive layer, auxiliary BPR] -> Adam
  -> AverageMeter update + step counters -- all libb200rec kernels on one stream, no host round trip.
On the C1-C3 shapes one SpMM layer is ~10 us of memory work, so per-kernel launch latency would dominate an eager
loop; the graph removes it.  The Adam moments are the tensors inside the trainer's torch.optim.Adam state, so
optimiser state_dicts stay interchangeable with the eager path.
"""
import os

import torch

from . import ops


def _adam_state(opt, p):
    """create (or fetch) torch.optim.Adam's per-parameter state in its own layout"""
    st = opt.state[p]
    if len(st) == 0:
        st['step'] = torch.tensor(0.0, dtype=torch.float32)
        st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
    return st


class BprEngine:
    def __init__(self, model, dataset, opt, batch_size, l2_reg, aux_reg=0.0, aux_dataset=None, seed=2021,
                 use_graph=True, partition=None, dim_shard=None, contrastive_reg=0.0):
        self.model, self.dataset, self.opt = model, dataset, opt
        self.kind = type(model).__name__
        if self.kind not in ('LightGCN', 'IGCN', 'IMF', 'MF', 'SGL', 'HALF'):
            raise ValueError('BprEngine supports LightGCN / IGCN / IMF / MF / SGL / HALF, got ' + self.kind)
        self.contrastive_reg = float(contrastive_reg)
        if self.kind in ('SGL', 'HALF') and (partition is not None or dim_shard is not None
                                             or getattr(model, '_dim_shard', None) is not None):
            raise ValueError('the contrastive step is single-GPU (its B x B logits couple every column and row)')
        if type(opt).__name__ != 'Adam':
            raise ValueError('the fused step implements Adam; use the autograd path for other optimisers')
        g = opt.param_groups[0]
        if g.get('weight_decay', 0) != 0 or g.get('amsgrad', False) or g.get('maximize', False):
            raise ValueError('fused Adam supports the default torch.optim.Adam options only')
        self.lr, (self.b1, self.b2), self.eps = g['lr'], g['betas'], g['eps']
        self.B, self.l2_reg, self.aux_reg, self.seed = batch_size, float(l2_reg), float(aux_reg), seed
        self.partition = partition
        self.shard = dim_shard if dim_shard is not None else getattr(model, '_dim_shard', None)
        if self.shard is not None and self.shard.world == 1:
            self.shard = None
        # both at once = the hybrid layout (DimShard.row_partition): column shards x row-partitioned replicas
        dev = model.device
        self.dev = dev
        D = model.embedding_size
        self.D = D
        i64 = dict(dtype=torch.int64, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.batch = torch.zeros((batch_size, 3), **i64)
        self.sample_step = torch.zeros(1, **i64)
        self.adam_step = torch.zeros(1, **i64)
        self.loss = torch.zeros(1, **f32)
        self.loss_accum = torch.zeros(2, dtype=torch.float64, device=dev)
        self.scratch = ops.bpr_scratch(batch_size, D, dev)
        self.dots = torch.zeros((batch_size, 3), **f32)  # per-sample (pos, neg, l2) partials when the dimension is sharded
        # ordered scatter (b200rec_bpr_group_rows): one store per distinct row, contributions summed in slot order
        self.ordered = 3 * batch_size <= ops.GROUP_CAP and os.environ.get('B200REC_ATOMIC_SCATTER', '0') != '1'
        self.grouping = ops.bpr_grouping(batch_size, dev) if self.ordered else None
        self.aux_grouping = None
        self.user_ptr, self.user_items = dataset.csr('train', device=dev)
        n = model.n_users + model.n_items
        self.n = n
        if self.kind == 'MF':
            self.table = model._joint
            self.params = [(model.user_embedding.weight, slice(0, model.n_users)),
                           (model.item_embedding.weight, slice(model.n_users, n))]
            self.grad = torch.zeros_like(self.table)
            self.m, self.v = torch.zeros_like(self.table), torch.zeros_like(self.table)
            for p, sl in self.params:
                st = _adam_state(opt, p)
                self.m[sl].copy_(st['exp_avg'])
                self.v[sl].copy_(st['exp_avg_sq'])
                st['exp_avg'], st['exp_avg_sq'] = self.m[sl], self.v[sl]
                p.grad = self.grad[sl]
                self.adam_step.fill_(int(st['step']))
        else:
            emb = model.embedding.weight
            if partition is not None and hasattr(partition, 'table'):
                # peer-memory row partition: the parameter table itself lives in memory the other ranks can store into
                partition.table.copy_(emb.data)
                emb.data = partition.table
            self.table = emb.data
            self.grad = torch.zeros_like(self.table)
            emb.grad = self.grad
            st = _adam_state(opt, emb)
            st['exp_avg'], st['exp_avg_sq'] = st['exp_avg'].to(dev), st['exp_avg_sq'].to(dev)
            self.m, self.v = st['exp_avg'], st['exp_avg_sq']
            self.adam_step.fill_(int(st['step']))
            self.g_rep = torch.zeros((n, D), **f32)
            self.row_flags = torch.zeros(n, dtype=torch.uint8, device=dev)
            # The two masked layers (last forward layer on the sampled rows, first backward hop from the sampled rows) can be
            # given their own hub-row chunk (B200REC_SPARSE_CHUNK).  Measured on C2 (ms/step): same plan 0.377, chunk 128
            # 0.387, 64 0.404, 32 0.435, 1024 0.494 -- the auto chunk is the optimum for them too, so the default is off.
            sc = int(os.environ.get('B200REC_SPARSE_CHUNK', '0'))
            self.adj_sparse = model.norm_adj.replan(sc) if (partition is None and sc > 0 and sc != model.norm_adj.chunk) \
                else model.norm_adj
            L = model.n_layers
            # rows within one hop of the batch (ops.mark_reach): forward layer L-1 is needed only there, and after the first
            # backward hop the gradient is non-zero only there.  The user half stays all-ones (see _reach_gain).
            self.reach = None
            reach_env = os.environ.get('B200REC_REACH', 'auto')
            if self.kind in ('LightGCN', 'IGCN') and L >= 2 and reach_env != '0' \
                    and (reach_env == '1' or self._reach_gain() >= 0.25):
                self.reach = torch.ones(n, dtype=torch.uint8, device=dev)
            # compacted work list of the last forward layer (only the sampled rows are needed): built beside the first layers
            self.live = None
            if partition is None and self.kind == 'LightGCN' and L >= 2 \
                    and os.environ.get('B200REC_NO_LIVE_LIST', '0') != '1':
                op = self.adj_sparse
                n_hub_items = int((op.item_dst[:op.n_items] < 0).sum())
                self.live = (torch.empty(max(op.n_items, 1), dtype=torch.int32, device=dev),
                             torch.zeros(1, dtype=torch.int32, device=dev), min(op.n_items, 3 * batch_size + n_hub_items))
            if partition is not None and hasattr(partition, 'rep'):
                self.rep, self.bufs = partition.rep, [partition.buf0, partition.buf1]
            else:
                self.rep = torch.empty((n, D), **f32)
                self.bufs = [torch.empty((n, D), **f32) if L >= 2 + i else None for i in range(2)]
            if self.kind in ('SGL', 'HALF'):
                nv = model.n_views
                self.user_flags = torch.zeros(n, dtype=torch.uint8, device=dev)
                self.view_rep = [torch.empty((n, D), **f32) for _ in range(nv)]
                self.view_g = [torch.zeros((n, D), **f32) for _ in range(nv)]
                self.view_grad = torch.empty((n, D), **f32)
                self.nce_ws = ops.infonce_workspace(batch_size, D, dev)
            if self.kind in ('IGCN', 'IMF'):
                self.x0 = torch.empty((n, D), **f32)
                self.dx0 = torch.empty((n, D), **f32)
                self.keep_bits = torch.zeros((model.feat_mat.nnz + 31) // 32, dtype=torch.int32, device=dev)
                self.row_scale = model.row_scale.clone()
                self.feat_fwd = model.feat_mat.fwd.with_scales(row_scale=self.row_scale)
                self.feat_bwd = model.feat_mat.bwd.with_scales(nbr_scale=self.row_scale)
                self.aux_batch = torch.zeros((batch_size, 3), **i64)
                self.g_w = torch.zeros(D, **f32)
                model.w.grad = self.g_w
                stw = _adam_state(opt, model.w)
                stw['exp_avg'], stw['exp_avg_sq'] = stw['exp_avg'].to(dev), stw['exp_avg_sq'].to(dev)
                self.m_w, self.v_w = stw['exp_avg'], stw['exp_avg_sq']
                self.aux = aux_dataset
                if aux_dataset is not None:
                    self.aux_ptr, self.aux_items = aux_dataset.csr('train', device=dev)
                    self.aux_grouping = ops.bpr_grouping(batch_size, dev) if self.ordered else None
        # out / addend traffic of the layers limited to the sampled rows (b200rec_spmm_f32_sel); B200REC_SEL=0 turns it off
        self.sel = os.environ.get('B200REC_SEL', '1') != '0'
        self._batch_early = True
        self.use_graph = use_graph and os.environ.get('B200REC_NO_GRAPH', '0') != '1'
        self._adj_ref = getattr(model, 'norm_adj', None)  # the step (and its captured graph) is built on THIS operand
        self._side_stream = torch.cuda.Stream(device=dev)
        fork_env = os.environ.get('B200REC_FORK', 'auto')  # auto: fork the batch prelude only when the tables fit L2
        self.fork = (self.table.numel() * 4 <= 96 * 2 ** 20) if fork_env == 'auto' else fork_env == '1'
        self._graphs = {}
        self._kernels = {}
        self.steps_done = 0

    def _reach_gain(self):
        """Share of the item-side edges that lies OUTSIDE one hop of an average batch: an item of degree d is missed by B
        uniformly drawn users with probability (1 - d/U)^B, so the share is sum_i d_i (1 - d_i/U)^B / E.  C4 (2 M users,
        batch 2048): 0.58 -- the two restricted layers skip that share of one side's gathers; C1-C3: < 0.05 (a batch's
        users reach nearly every item that carries edges), where the extra mask costs more than it saves."""
        from .graph import reach_gain
        m = self.model
        return reach_gain(torch.bincount(self.user_items.long(), minlength=m.n_items), m.n_users, self.B)

    # ------------------------------------------------------------------------------------------------ one step
    def _propagate_fwd(self, x0, join=None, adj=None, flags=None, rep=None):
        """L x SpMM + layer mean into self.rep.  Only the LAST layer needs the batch (it is restricted to the sampled
        rows), so `join` -- which makes the main stream wait for the side stream that draws the batch and marks its rows --
        is called right before it: sampling overlaps the first L-1 layers inside the step graph."""
        m = self.model
        L = m.n_layers
        own = adj is None  # the model's full graph; otherwise an augmented view (SGL / HALF)
        adj = m.norm_adj if own else adj
        reach = self.reach if (own and flags is None) else None  # the batch's one-hop mask goes with the batch's row mask
        flags = self.row_flags if flags is None else flags
        rep = self.rep if rep is None else rep
        if self.partition is not None:
            if join:
                join()
            self.partition.propagate_fwd(adj, x0, L, self.bufs, rep, needed_rows=flags, reach_rows=reach)
            return
        if L == 0:
            if join:
                join()
            ops.propagate_fwd(adj, x0, L, self.bufs, rep)
            return
        inv = 1.0 / (L + 1)
        src = x0
        for k in range(L):  # the layer loop of b200rec_propagate_fwd, opened up for the join and the sparse plan
            last = k == L - 1
            near = reach is not None and k == L - 2  # layer L-1 feeds the last layer only: needed within one hop of the batch
            if join and (near or (last and (reach is None or L < 2))):
                join()
            y = None if last else self.bufs[k & 1]
            if last and own and join is not None and self.live is not None and flags is self.row_flags:
                ops.spmm_live(self.adj_sparse, src, self.live[0], self.live[1], self.live[2], addend=x0 if k == 0 else rep,
                              out=rep, out_scale=inv)
            else:
                # the step reads rep only at the flagged rows: the running layer sum is kept only there (when the batch is
                # known before the first layer, i.e. the prelude is not forked onto the side stream)
                keep = flags if (self.sel and not last and (join is None or self._batch_early)) else None
                ops.spmm(self.adj_sparse if (last and own) else adj, src, y=y, addend=x0 if k == 0 else rep, out=rep,
                         out_scale=inv if last else 1.0, dst_flags=flags if last else (reach if near else None),
                         out_rows=keep, out_mode=1)
            src = y

    def _propagate_bwd(self, out, adj=None, flags=None, g=None):
        m = self.model
        own = adj is None
        adj = m.norm_adj if own else adj
        reach = self.reach if (own and flags is None) else None
        flags = self.row_flags if flags is None else flags
        g = self.g_rep if g is None else g
        if self.partition is not None:
            self.partition.propagate_bwd(adj, g, m.n_layers, self.bufs, out, nonzero_rows=flags, reach_rows=reach)
            return
        L = m.n_layers
        if L == 0:
            ops.propagate_bwd(adj, g, L, self.bufs, out)
            return
        inv = 1.0 / (L + 1)
        src = g
        for k in range(1, L + 1):  # H_k = G + A H_{k-1}; the first hop reads only the <= 3B non-zero rows of G
            last = k == L
            dst = out if last else self.bufs[(k - 1) & 1]
            # H_1 is zero outside one hop of the batch: those rows are neither computed (unless H_1 is the result) nor
            # gathered by the second hop
            ops.spmm(self.adj_sparse if (k == 1 and own) else adj, src, addend=g, out=dst, out_scale=inv if last else 1.0,
                     src_flags=flags if k == 1 else (reach if k == 2 else None),
                     dst_flags=reach if (k == 1 and not last) else None,
                     out_rows=flags if self.sel else None, out_mode=2)  # G is zero outside the flagged rows: not read there
            src = dst

    def _bpr(self, rep, batch, item_offset, l2_reg, reg_mode, g_rep, w=None, g_w=None, loss_scale=1.0, grouping=None,
             accumulate=False, emb0=None, l2_emb0=0.0):
        """grouping: the batch's (order, coef) from ops.bpr_group_rows -> deterministic aggregated
        scatter (g_rep rows are STORED unless accumulate); None -> one 128-bit red per sample and role"""
        if grouping is not None:
            if self.shard is None:
                ops.bpr_fwd_bwd_ordered(rep, batch, item_offset, l2_reg, reg_mode, g_rep, self.loss, self.scratch, grouping,
                                        w=w, g_w=g_w, loss_scale=loss_scale, accumulate=accumulate, emb0=emb0, l2_emb0=l2_emb0)
                return
            ops.bpr_fwd_bwd_ordered(rep, batch, item_offset, l2_reg, reg_mode, g_rep, self.loss, self.scratch, grouping,
                                    dots=self.dots, phase=1, w=w, g_w=g_w, loss_scale=loss_scale)
            self.shard.all_reduce_sum(self.dots)
            ops.bpr_fwd_bwd_ordered(rep, batch, item_offset, l2_reg, reg_mode, g_rep, self.loss, self.scratch, grouping,
                                    dots=self.dots, phase=2, loss_weight=1.0 if self.shard.rank == 0 else 0.0, w=w, g_w=g_w,
                                    loss_scale=loss_scale, accumulate=accumulate, emb0=emb0, l2_emb0=l2_emb0)
            return
        if self.shard is None:
            ops.bpr_fwd_bwd(rep, batch, item_offset, l2_reg, reg_mode, g_rep, self.loss, self.scratch, w=w, g_w=g_w,
                            loss_scale=loss_scale)
            return
        # embedding dimension sharded over the ranks: partial dot products -> all-reduce -> this rank's gradient columns
        ops.bpr_fwd_bwd_sharded(rep, batch, item_offset, l2_reg, reg_mode, g_rep, self.loss, self.scratch, self.dots, 1,
                                w=w, g_w=g_w, loss_scale=loss_scale)
        self.shard.all_reduce_sum(self.dots)
        ops.bpr_fwd_bwd_sharded(rep, batch, item_offset, l2_reg, reg_mode, g_rep, self.loss, self.scratch, self.dots, 2,
                                loss_weight=1.0 if self.shard.rank == 0 else 0.0, w=w, g_w=g_w, loss_scale=loss_scale)

    def _adam(self, param, grad, m, v):
        if self.partition is not None and param is self.table and hasattr(self.partition, 'adam'):
            self.partition.adam(param, grad, m, v, self.adam_step, self.lr, self.b1, self.b2, self.eps)
            return
        if self.partition is not None and param is self.table:
            for lo, hi in self.partition.my_blocks:  # each rank updates the rows it owns, then the blocks are exchanged
                ops.adam_step(param[lo:hi], grad[lo:hi], m[lo:hi], v[lo:hi], self.adam_step, self.lr, self.b1, self.b2, self.eps)
            self.partition.exchange(param)
            return
        ops.adam_step(param, grad, m, v, self.adam_step, self.lr, self.b1, self.b2, self.eps)

    def _body(self, sample, draw_mask=True):
        m, B = self.model, self.B
        nu = m.n_users
        main = torch.cuda.current_stream()
        # the batch's small kernels (sampler, row grouping, row marks, live work list) are needed by the last forward layer
        # only: `prelude` = they are issued ahead of the layers and the last layer runs on the compacted work list;
        # `forked` = they run on a side stream beside the first layers instead of in front of them.  Forking pays on the
        # L2-resident shapes (C2, same box: 0.386 ms/step forked, 0.403 serial) and not on the large ones, where ~60 us of
        # prelude is nothing and the forked graph measured WORSE with host batches (C4, same box: 76.2 ms device-sampled,
        # 90.5 ms with a host batch copied in before the graph; serial 75.6 / 75.7).  self.fork: tables fit L2.
        prelude = self.kind == 'LightGCN' and self.partition is None and m.n_layers >= 2
        forked = prelude and self.fork
        self._batch_early = not forked
        side = self._side_stream if forked else main
        grp = self.grouping
        # g_rep is all-zero between steps: with the ordered scatter the BPR kernel STORES the touched rows and they are
        # cleared again once the backward chain has consumed them (no [N, D] memset per step: 1.5 GB on C4)
        lazy_clear = grp is not None and self.kind in ('LightGCN', 'IGCN', 'IMF')
        if forked:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            if sample:
                ops.bpr_sample(self.user_ptr, self.user_items, self.dataset.n_users, self.dataset.n_items, self.seed,
                               self.sample_step, B, out=self.batch)
            if grp is not None:
                ops.bpr_group_rows(self.batch, nu, grp)  # full-grid rank sort, a few microseconds beside the first layers
            if self.kind != 'MF':
                # the step reads rep only at the <= 3B sampled rows, and G is non-zero only there: the last forward layer
                # and the first backward hop are restricted to them (bit-identical on the rows that matter)
                self.row_flags.zero_()
                ops.mark_rows(self.batch, nu, self.row_flags)
                if self.reach is not None:
                    self.reach[nu:].zero_()
                    ops.mark_reach(self.batch, nu, self.user_ptr, self.user_items, self.reach)
                if prelude and self.live is not None:
                    ops.live_items(self.adj_sparse, self.row_flags, self.live[0], self.live[1])
            if forked and not lazy_clear:
                self.g_rep.zero_()  # needed only by the BPR kernel: cleared beside the first forward layers
            if forked:
                self.loss.zero_()  # off the critical path too
        join = (lambda: main.wait_stream(side)) if forked else ((lambda: None) if prelude else None)
        if not forked:
            self.loss.zero_()
        if self.kind == 'MF':
            self.grad.zero_()
            self._bpr(self.table, self.batch, nu, self.l2_reg, 1, self.grad, grouping=grp)
            self._adam(self.table, self.grad, self.m, self.v)
        elif self.kind == 'LightGCN':
            if not forked and not lazy_clear:
                self.g_rep.zero_()
            self._propagate_fwd(self.table, join)
            fold = grp is not None and self.l2_reg != 0.0  # layer-0 L2: loss inside the BPR launch, gradient + row clear in one
            self._bpr(self.rep, self.batch, nu, 0.0, 0, self.g_rep, grouping=grp, emb0=self.table if fold else None,
                      l2_emb0=self.l2_reg if fold else 0.0)
            self._propagate_bwd(self.grad)
            if fold:
                ops.bpr_l2_emb0_ordered(self.table, self.batch, nu, self.l2_reg, self.grad, None, None, grp,
                                        clear_table=self.g_rep if lazy_clear else None)
            else:
                if lazy_clear:
                    ops.clear_rows(self.batch, nu, self.g_rep)
                if self.l2_reg != 0.0:
                    ops.bpr_l2_emb0(self.table, self.batch, nu, self.l2_reg, self.grad, self.loss, self.scratch)
            self._adam(self.table, self.grad, self.m, self.v)
        elif self.kind in ('SGL', 'HALF'):
            # SGLTrainer / HALFTrainer.train_one_epoch (trainer.py:440-456): BPR on the full graph + contrastive_reg *
            # InfoNCE between two views' rows of the batch users; the views need their last layer (and send gradient)
            # only at those user rows.  Three (two) propagations forward, three (two) Horner chains backward, summed.
            views = [m.norm_aug_adj1] + ([m.norm_aug_adj2] if m.n_views == 2 else [])
            self.user_flags.zero_()
            self.user_flags.index_fill_(0, self.batch[:, 0], 1)
            self.g_rep.zero_()
            for gv in self.view_g:
                gv.zero_()
            self._propagate_fwd(self.table)
            for adj, rep in zip(views, self.view_rep):
                self._propagate_fwd(self.table, adj=adj, flags=self.user_flags, rep=rep)
            self._bpr(self.rep, self.batch, nu, self.l2_reg, 1 if self.l2_reg != 0.0 else 0, self.g_rep, grouping=grp)
            if self.kind == 'SGL':
                q, k, gq, gk = self.view_rep[0], self.view_rep[1], self.view_g[0], self.view_g[1]
            else:
                q, k, gq, gk = self.rep, self.view_rep[0], self.g_rep, self.view_g[0]
            ops.infonce_fwd_bwd(q, k, self.loss, gq, gk, self.nce_ws, temperature=m.temperature,
                                loss_scale=self.contrastive_reg, rows=self.batch, row_stride=3, n=B)
            self._propagate_bwd(self.grad)
            for adj, gv in zip(views, self.view_g):
                self._propagate_bwd(self.view_grad, adj=adj, flags=self.user_flags, g=gv)
                self.grad.add_(self.view_grad)
            self._adam(self.table, self.grad, self.m, self.v)
        else:  # IGCN / IMF
            p = float(m.dropout)
            keep, inv_keep = None, 1.0
            if p > 0.0:
                if draw_mask:
                    ops.dropout_bits(m.feat_mat.nnz, p, m.dropout_seed, self.sample_step, out=self.keep_bits)
                keep, inv_keep = self.keep_bits, 1.0 / (1.0 - p)
            if self.aux is not None and sample:
                ops.bpr_sample(self.aux_ptr, self.aux_items, self.aux.n_users, self.aux.n_items, self.seed + 1,
                               self.sample_step, B, out=self.aux_batch)
            if not lazy_clear:
                self.g_rep.zero_()
            ops.spmm(self.feat_fwd, self.table, keep_bits=keep, post_scale=inv_keep, y=self.x0)
            self._propagate_fwd(self.x0)
            self._bpr(self.rep, self.batch, nu, self.l2_reg, 1 if self.l2_reg != 0.0 else 0, self.g_rep, grouping=grp)
            self._propagate_bwd(self.dx0)
            if lazy_clear:
                ops.clear_rows(self.batch, nu, self.g_rep)
            ops.spmm(self.feat_bwd, self.dx0, keep_bits=keep, post_scale=inv_keep, y=self.grad)
            if self.aux is not None:
                tu = len(m.user_map)
                if self.aux_grouping is not None:
                    ops.bpr_group_rows(self.aux_batch, tu, self.aux_grouping)
                self._bpr(self.table, self.aux_batch, tu, 0.0, 0, self.grad, w=m.w.data, g_w=self.g_w,
                          loss_scale=self.aux_reg, grouping=self.aux_grouping, accumulate=True)
                self._adam(m.w.data, self.g_w, self.m_w, self.v_w)
            self._adam(self.table, self.grad, self.m, self.v)
        ops.step_advance(self.adam_step, self.sample_step, self.loss, self.loss_accum, B)

    def _run(self, sample, draw_mask=True):
        if getattr(self.model, 'norm_adj', None) is not self._adj_ref:
            raise RuntimeError('model.norm_adj was replaced after the training engine was built: create a new trainer')
        if not self.use_graph:
            self._body(sample, draw_mask)
            return
        ver = getattr(self.model, 'aug_version', 0)  # SGL / HALF redraw their views every epoch: new operands, new graph
        if ver != getattr(self, '_graph_ver', 0):
            torch.cuda.synchronize()
            self._graphs.clear()
            self._graph_ver = ver
        g = self._graphs.get((sample, draw_mask))
        if g is None:
            # warm-up on a side stream (lazy module loading, cudaFuncSetAttribute) with state restored afterwards
            snap = [t.clone() for t in self._state_tensors()]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            from . import _abi
            n0 = _abi.launch_count()
            with torch.cuda.stream(s):
                self._body(sample, draw_mask)
            self._kernels[(sample, draw_mask)] = _abi.launch_count() - n0
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            for t, c in zip(self._state_tensors(), snap):
                t.copy_(c)
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread may touch CUDA while a sharded step (with its all-reduce) is captured
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._body(sample, draw_mask)
            for t, c in zip(self._state_tensors(), snap):  # capture does not execute, but keep state exact anyway
                t.copy_(c)
            self._graphs[(sample, draw_mask)] = g
        g.replay()

    def _state_tensors(self):
        ts = [self.table, self.m, self.v, self.adam_step, self.sample_step, self.loss_accum, self.loss]
        if self.kind in ('IGCN', 'IMF'):
            ts += [self.model.w.data, self.m_w, self.v_w]
        return ts

    # ------------------------------------------------------------------------------------------------ public
    def step(self, host_batch=None, host_aux_batch=None, host_keep_bits=None):
        """One optimiser step.  host_batch: optional pinned int64 [B,3] (the reference's DataLoader batch,
        trainer.py:414); without it the triples are drawn on device."""
        if host_batch is not None:
            self.batch.copy_(host_batch.reshape(self.B, 3), non_blocking=True)
            if host_aux_batch is not None:
                self.aux_batch.copy_(host_aux_batch.reshape(self.B, 3), non_blocking=True)
            draw_mask = True
            if host_keep_bits is not None:
                self.keep_bits.copy_(host_keep_bits, non_blocking=True)
                draw_mask = False
            elif self.kind in ('IGCN', 'IMF') and self.model.dropout > 0 and self.model.dropout_rng == 'host':
                self.keep_bits.copy_(self.model._keep_bits())  # the reference's CPU torch.rand draw (model.py:4020)
                draw_mask = False
            self._run(False, draw_mask)
        else:
            self._run(True)
        self.model._rep_cache = None  # kernels update parameters in place, behind autograd's version counter
        self.steps_done += 1

    def kernels_per_step(self, sample=True, draw_mask=True):
        """libb200rec kernel launches inside one step (counted while the step was warmed up for capture)"""
        return self._kernels.get((sample, draw_mask))

    def close(self):
        """drop the captured graphs (needed before torch.distributed.destroy_process_group() when a graph holds
        NCCL kernels: tearing the communicator down under a live graph blocks)"""
        torch.cuda.synchronize()
        self._graphs.clear()

    def refresh_row_scale(self):
        """after IGCN.feat_mat_anneal(): the graph reads the scale vector in place"""
        if self.kind in ('IGCN', 'IMF'):
            self.row_scale.copy_(self.model.row_scale)

    def reset_meter(self):
        self.loss_accum.zero_()

    def meter_avg(self):
        a = self.loss_accum.clone()
        if self.shard is not None:  # shard rank 0 carries the BPR term, every rank its own columns' L2 term
            self.shard.all_reduce_sum(a[:1])
        a = a.cpu()
        return float(a[0] / a[1]) if float(a[1]) > 0 else 0.0

    def last_loss(self):
        if self.shard is not None:
            return float(self.shard.all_reduce_sum(self.loss.clone()).item())
        return float(self.loss.item())

    def note_external_step(self, loss, n_batch):
        """an optimiser step was taken outside the engine (eager autograd + torch.optim.Adam on the shared moments):
        advance the device step count and the loss meter as a replay would have"""
        self.adam_step += 1
        self.loss_accum += torch.tensor([float(loss) * n_batch, float(n_batch)], dtype=torch.float64, device=self.dev)
        self.model._rep_cache = None

    def sync_optimizer_state(self):
        """write the step count back into torch.optim.Adam's state (it keeps `step` as a CPU scalar tensor)"""
        t = float(self.adam_step.item())
        for p in [q for grp in self.opt.param_groups for q in grp['params']]:
            if p in self.opt.state and 'step' in self.opt.state[p]:
                self.opt.state[p]['step'] = torch.tensor(t, dtype=torch.float32)
        self.model._rep_cache = None
