"""Multi-GPU decompositions of the hot path (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no multi-device code.  Two decompositions are built, both bit-identical
to the single-GPU result on every output element:

* `RowPartition` -- the scheme BASELINE.json names: contiguous row blocks of the [users | items] index space,
  balanced by nnz; rank p computes rows [r_p, r_{p+1}) of every propagated table and the blocks are all-gathered
  after each layer (backward: the same, A is symmetric).  Per-row sums are untouched by the partition.  Exchange volume
  per layer is (P-1)/P * N * D * 4 bytes per rank, which is communication-bound for P >= 4.

* `DimShard` -- shard the embedding DIMENSION instead: every rank holds D/P columns of every [N, D] table.  Propagation
  (X <- A X), the layer mean, the inductive layer, all gradients and Adam act on each column independently, so the
  L forward and L backward layers need NO exchange at all; the only coupling is the BPR score <u, i> = sum over
  columns, one all-reduce of a [B, 3] fp32 tensor (24 KB at B = 2048) per step.  Evaluation all-gathers the columns of
  the final representation once and shards the USERS of the score / top-K sweep.  This is the default for training.
"""
import numpy as np
import torch
import torch.distributed as dist


class DimShard:
    """rank r of `world` holds columns [cols(d)] of every table.  Shards narrower than `min_cols` columns are not worth
    having (the SpMM becomes per-edge-overhead-bound: C2 one layer 57 us at 64 columns, 42 at 32, 40 at 16, 50 at 8), so
    the column split stops at P' = the largest power of two <= world with D / P' >= min_cols, and the world is arranged
    as world / P' identical replicas of a P'-way shard group (same seed, same sampler stream -> same numbers, no traffic
    between replicas).  Evaluation always shards the USERS over the whole world."""

    def __init__(self, rank=0, world=1, group=None, min_cols=None):
        import os
        self.world_rank, self.world_size, self.world_group = rank, world, group
        env = os.environ.get("B200REC_MIN_COLS")
        self.min_cols = min_cols if min_cols is not None else (int(env) if env else None)  # None: chosen in configure()
        self.rank, self.world, self.group = rank, world, group  # shard-group coordinates, fixed by configure()
        self._configured = world == 1

    def configure(self, d, n_rows=None):
        """collective (every rank of the world calls it with the same d): pick P' and create the shard groups.
        Default narrowest shard: 16 columns when the table streams from HBM (C4: 75.6 -> 13.8 ms/step on 8 GPUs), 32 when
        it is L2-resident -- there a 16-column layer is no faster than a 32-column one (40 vs 42 us on C2) and the wider
        all-reduce group only adds latency (C2 on 4 GPUs: 0.409 ms/step as 4 x 16 columns, 0.316 as 2 x 32 x 2 replicas)."""
        if self._configured:
            return self
        if self.min_cols is None:
            # 32 columns (128-byte rows) is the narrowest shard worth having on either kind of table: L2-resident ones
            # gain nothing below it (C2: 42 us at 32 columns, 40 at 16), HBM-streamed ones gather 64-byte rows at 6.5 TB/s
            # against 8.8 TB/s for 128-byte rows (C4, round 2) -- there the remaining factor goes to the rows
            # (row_partition), for L2-resident tables to identical replicas
            self.min_cols = 32
        p = 1
        while p * 2 <= self.world_size and d % (p * 2) == 0 and d // (p * 2) >= self.min_cols:
            p *= 2
        self.world = p
        self.rank = self.world_rank % p
        self.replica = self.world_rank // p
        self.n_replicas = self.world_size // p
        self.group = self.world_group
        self.peer_group = None  # the ranks that hold the SAME columns in the other replicas (hybrid row partition)
        if p != self.world_size:
            for rep in range(self.world_size // p):  # every rank creates every group, in the same order
                g = dist.new_group(list(range(rep * p, (rep + 1) * p)))
                if rep == self.replica:
                    self.group = g
            for col in range(p):
                g = dist.new_group(list(range(col, self.world_size, p)))
                if col == self.rank:
                    self.peer_group = g
        self._configured = True
        return self

    def row_partition(self, adj, d, kind="peer"):
        """Hybrid layout for tables that stream from HBM: the world is (world // P') replicas of a P'-way column shard;
        instead of running identical replicas, the replicas that hold the same columns split the ROWS among them
        (PeerRowPartition over `peer_group`: finished rows are pushed into the peers' tables over NVLink inside the SpMM /
        Adam epilogues).  Why: the gather rate falls with the row width -- on C4 a 16-column shard (64-byte rows) gathers at
        6.5 TB/s, a 32-column one at 8.8 -- so 8 GPUs as 4 column shards x 2 row halves move the same bytes per GPU faster
        than 8 column shards, and the index scan per GPU halves.  Returns None when there is one replica."""
        if getattr(self, "n_replicas", 1) <= 1:
            return None
        cls = PeerRowPartition if kind == "peer" else RowPartition
        if cls is PeerRowPartition:
            return PeerRowPartition(adj, self.replica, self.n_replicas, d, group=self.peer_group)
        return RowPartition(adj, self.replica, self.n_replicas, group=self.peer_group)

    def cols(self, d):
        assert d % self.world == 0, "embedding_size must be divisible by the shard count"
        w = d // self.world
        return self.rank * w, (self.rank + 1) * w

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def gather_cols(self, local):
        """[N, D/P'] on every rank of a shard group -> [N, D] on every rank"""
        if self.world == 1:
            return local
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather(parts, local.contiguous(), group=self.group)
        return torch.cat(parts, dim=1)

    def user_range(self, n_users):
        per = (n_users + self.world_size - 1) // self.world_size
        return min(n_users, self.world_rank * per), min(n_users, (self.world_rank + 1) * per)

    def gather_user_rows(self, local, n_users):
        """rows of the rank's user_range -> all rows on every rank (user-sharded evaluation over the whole world)"""
        if self.world_size == 1:
            return local
        per = (n_users + self.world_size - 1) // self.world_size
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        parts = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(parts, pad, group=self.world_group)
        return torch.cat(parts, dim=0)[:n_users]


def shard_model_dims(model, shard):
    """Keep only this rank's columns of every embedding parameter (call before the optimiser is created).
    Every rank must have built the model from the same seed, so the slices are slices of ONE initialisation."""
    import torch.nn as nn
    d = model.embedding_size
    shard.configure(d, n_rows=model.n_users + model.n_items)
    if shard.world == 1:
        model._dim_shard = shard
        return model
    lo, hi = shard.cols(d)
    if type(model).__name__ == 'MF':
        joint = model._joint[:, lo:hi].contiguous()
        model._joint = joint
        model.user_embedding.weight = nn.Parameter(joint[:model.n_users])
        model.item_embedding.weight = nn.Parameter(joint[model.n_users:])
        model.user_embedding.embedding_dim = model.item_embedding.embedding_dim = hi - lo
    else:
        model.embedding.weight = nn.Parameter(model.embedding.weight.data[:, lo:hi].contiguous())
        model.embedding.embedding_dim = hi - lo
        if hasattr(model, 'w'):
            model.w = nn.Parameter(model.w.data[lo:hi].contiguous())
    model.full_embedding_size = d
    model.embedding_size = hi - lo
    model._dim_shard = shard
    model._rep_cache = None
    return model


def nnz_balanced_bounds(rowptr, world):
    """split points r_0 = 0 <= r_1 <= ... <= r_P = n_rows with about nnz/P entries per block"""
    rp = np.asarray(rowptr, dtype=np.int64)
    n = len(rp) - 1
    targets = (np.arange(1, world) * rp[-1]) // world
    cuts = np.searchsorted(rp, targets, side='left')
    return [0] + [int(min(max(c, 0), n)) for c in cuts] + [n]


class RowPartition:
    """Row-partitioned propagation with a per-layer exchange of the row blocks.

    Each rank owns one block of USER rows and one block of ITEM rows (both nnz-balanced over the ranks) rather than one
    contiguous range of the joint index space: user rows gather from the item table and item rows from the (larger,
    colder) user table, so a single range would give the ranks unequal cache behaviour (measured on C4, 2 GPUs: one rank
    all users, the other all items -> 54.6 ms/step against 75.6 on one GPU)."""

    def __init__(self, adj, rank, world, group=None, bounds=None):
        self.rank, self.world, self.group = rank, world, group
        rp = adj.rowptr.cpu().numpy().astype(np.int64)
        nu = int(getattr(adj, "phase_split", 0))
        if bounds is not None:
            self.blocks = [[(bounds[r], bounds[r + 1])] for r in range(world)]
        elif 0 < nu < adj.n_rows:
            ub = nnz_balanced_bounds(rp[: nu + 1], world)
            ib = nnz_balanced_bounds(rp[nu:] - rp[nu], world)
            self.blocks = [[(ub[r], ub[r + 1]), (nu + ib[r], nu + ib[r + 1])] for r in range(world)]
        else:
            b = nnz_balanced_bounds(rp, world)
            self.blocks = [[(b[r], b[r + 1])] for r in range(world)]
        self.bounds = [blk[0][0] for blk in self.blocks] + [self.blocks[-1][0][1]]  # first-block bounds (diagnostics)
        self.my_blocks = [(lo, hi) for lo, hi in self.blocks[rank] if hi > lo]
        self.lo, self.hi = self.blocks[rank][0]
        self.n_local_rows = sum(hi - lo for lo, hi in self.my_blocks)
        self.local_op = adj.row_slice(self.blocks[rank])

    def exchange(self, buf):
        """in place: after the call every rank holds every row block of `buf` ([n_rows, D])"""
        if self.world == 1:
            return
        works = []
        for p in range(self.world):
            src = dist.get_global_rank(self.group, p) if self.group is not None else p  # partition rank -> world rank
            for lo, hi in self.blocks[p]:
                if hi > lo:
                    works.append(dist.broadcast(buf[lo:hi], src=src, group=self.group, async_op=True))
        for w in works:
            w.wait()

    def propagate_fwd(self, adj, x0, n_layers, bufs, mean_out, needed_rows=None, reach_rows=None):
        """same contract as ops.propagate_fwd; x0 must be complete on every rank, mean_out is complete on return.
        reach_rows: optional byte mask of the rows within one hop of needed_rows (ops.mark_reach) -- layer L-1 is computed
        only there"""
        from . import ops
        if n_layers == 0:
            mean_out.copy_(x0)
            return
        inv = 1.0 / (n_layers + 1)
        src = x0
        for k in range(n_layers):
            last = k == n_layers - 1
            y = None if last else bufs[k & 1]
            # the caller reads mean_out only at needed_rows: the running layer sum is kept only there (out_mode 1)
            ops.spmm(self.local_op, src, y=y, addend=x0 if k == 0 else mean_out, out=mean_out,
                     out_scale=inv if last else 1.0,
                     dst_flags=needed_rows if last else (reach_rows if k == n_layers - 2 else None),
                     out_rows=None if last else needed_rows, out_mode=1)
            if not last:
                self.exchange(y)
                src = y
        self.exchange(mean_out)

    def propagate_bwd(self, adj, g, n_layers, bufs, dx0, nonzero_rows=None, reach_rows=None):
        """dx0 rows of this rank's blocks are valid on return (each rank updates only the parameters it owns)"""
        from . import ops
        if n_layers == 0:
            dx0.copy_(g)
            return
        inv = 1.0 / (n_layers + 1)
        src = g
        for k in range(1, n_layers + 1):
            last = k == n_layers
            dst = dx0 if last else bufs[(k - 1) & 1]
            ops.spmm(self.local_op, src, addend=g, out=dst, out_scale=inv if last else 1.0,
                     src_flags=nonzero_rows if k == 1 else (reach_rows if k == 2 else None),
                     dst_flags=reach_rows if (k == 1 and not last) else None, out_rows=nonzero_rows, out_mode=2)
            if not last:
                self.exchange(dst)
                src = dst


# ------------------------------------------------------------------------------------------------ peer memory (CUDA IPC)
class _RawCuda:
    """__cuda_array_interface__ view of a raw device pointer (so torch can wrap memory this library allocated)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _wrap(ptr, nbytes, device):
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)


class PeerTable:
    """one [rows, d] fp32 (or raw byte) buffer per rank, each mapped into every other rank of the node"""

    def __init__(self, nbytes, rank, world, group, device):
        import ctypes as C
        from . import _abi
        lib = _abi.load()
        ptr = C.c_void_p(0)
        handle = (C.c_uint8 * 64)()
        _abi.check(lib.b200rec_peer_alloc(nbytes, C.byref(ptr), handle), "peer_alloc")
        self.ptr, self.nbytes, self.rank, self.world = ptr.value, nbytes, rank, world
        self.bytes = _wrap(self.ptr, nbytes, device)
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=group)
        self.peer_ptrs = []
        for r in range(world):
            if r == rank:
                self.peer_ptrs.append(self.ptr)
                continue
            h = (C.c_uint8 * 64)(*handles[r].cpu().tolist())
            p = C.c_void_p(0)
            _abi.check(lib.b200rec_peer_open(h, C.byref(p)), "peer_open")
            self.peer_ptrs.append(p.value)

    def others(self, byte_offset=0):
        """HOST ctypes array of the other ranks' pointers (+ offset), as the *_peer entry points take them"""
        import ctypes as C
        ptrs = [p + byte_offset for r, p in enumerate(self.peer_ptrs) if r != self.rank]
        return (C.c_void_p * max(len(ptrs), 1))(*ptrs)


class PeerRowPartition(RowPartition):
    """RowPartition whose per-layer exchange is fused into the producing kernels: the SpMM epilogue (and the Adam
    update) store every finished row into all peers' tables over NVLink while the kernel is still gathering for its next
    rows; what remains per layer is a two-kernel signal / wait hand-shake.  No NCCL on the data path."""

    def __init__(self, adj, rank, world, d, group=None):
        super().__init__(adj, rank, world, group)
        import ctypes as C
        dev = adj.device
        n = adj.n_rows
        self.d, self.n_rows = d, n
        nb = n * d * 4
        self._tables = {k: PeerTable(nb, rank, world, group, dev) for k in ("table", "buf0", "buf1", "rep")}
        self.table, self.buf0, self.buf1, self.rep = (self._tables[k].bytes.view(torch.float32).view(n, d)
                                                      for k in ("table", "buf0", "buf1", "rep"))
        self._flags = PeerTable(256, rank, world, group, dev)
        self.flag_ptrs = torch.tensor(self._flags.peer_ptrs, dtype=torch.int64, device=dev)  # device array of pointers
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.n_others = world - 1
        self._C = C
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def _others(self, name):
        return self._tables[name].others()

    def _name_of(self, t):
        for k in ("table", "buf0", "buf1", "rep"):
            if t.data_ptr() == getattr(self, k).data_ptr():
                return k
        raise ValueError("tensor is not one of this partition's peer tables")

    def handshake(self):
        """everyone's pushed rows have landed AND everyone is done reading the tables of the previous phase"""
        from . import _abi
        from ._abi import check, ptr, stream_ptr
        lib = _abi.load()
        check(lib.b200rec_peer_signal(ptr(self.flag_ptrs), self.world, self.rank, ptr(self.epoch), 1, stream_ptr()), "peer_signal")
        check(lib.b200rec_peer_wait(self._C.c_void_p(self._flags.ptr), self.world, ptr(self.epoch), 1, stream_ptr()), "peer_wait")

    def exchange(self, buf):
        self.handshake()  # the rows were pushed by the kernel that produced them

    def _spmm(self, src, y=None, addend=None, out=None, out_scale=1.0, dst_flags=None, src_flags=None, push_y=False,
              push_out=False, out_rows=None, out_mode=0):
        from . import _abi
        from ._abi import check, ptr, stream_ptr
        C = self._C
        py = self._others(self._name_of(y)) if push_y else None
        po = self._others(self._name_of(out)) if push_out else None
        check(_abi.load().b200rec_spmm_f32_peer(C.byref(self.local_op.struct()), ptr(src), src.shape[1], None, 1.0, ptr(y),
                                                ptr(addend), ptr(out), out_scale, ptr(dst_flags), ptr(src_flags),
                                                self.n_others, py, po, ptr(out_rows), out_mode if out_rows is not None else 0,
                                                stream_ptr()), "spmm_f32_peer")

    def propagate_fwd(self, adj, x0, n_layers, bufs, mean_out, needed_rows=None, reach_rows=None):
        assert mean_out is self.rep and x0.shape[1] == self.d
        if n_layers == 0:
            mean_out.copy_(x0)
            return
        inv = 1.0 / (n_layers + 1)
        bufs = [self.buf0, self.buf1]
        src = x0
        for k in range(n_layers):
            last = k == n_layers - 1
            y = None if last else bufs[k & 1]
            self._spmm(src, y=y, addend=x0 if k == 0 else mean_out, out=mean_out, out_scale=inv if last else 1.0,
                       dst_flags=needed_rows if last else (reach_rows if k == n_layers - 2 else None), push_y=not last,
                       push_out=last, out_rows=None if last else needed_rows, out_mode=1)
            self.handshake()
            src = y

    def propagate_bwd(self, adj, g, n_layers, bufs, dx0, nonzero_rows=None, reach_rows=None):
        if n_layers == 0:
            dx0.copy_(g)
            return
        inv = 1.0 / (n_layers + 1)
        bufs = [self.buf0, self.buf1]
        src = g
        for k in range(1, n_layers + 1):
            last = k == n_layers
            dst = dx0 if last else bufs[(k - 1) & 1]
            self._spmm(src, addend=g, out=dst, out_scale=inv if last else 1.0,
                       src_flags=nonzero_rows if k == 1 else (reach_rows if k == 2 else None),
                       dst_flags=reach_rows if (k == 1 and not last) else None, push_out=not last,
                       out_rows=nonzero_rows, out_mode=2)
            if not last:
                self.handshake()
                src = dst

    def adam(self, param, grad, m, v, step, lr, b1, b2, eps):
        """update the rows this rank owns and store them into every peer's table; hand-shakes on both sides"""
        from . import _abi
        from ._abi import check, ptr, stream_ptr
        assert param.data_ptr() == self.table.data_ptr()
        self.handshake()  # nobody still reads the old parameters (layer-0 L2 term, first forward layer)
        d = self.d
        for lo, hi in self.my_blocks:
            off = lo * d * 4
            check(_abi.load().b200rec_adam_step_peer(ptr(param[lo:hi]), ptr(grad[lo:hi]), ptr(m[lo:hi]), ptr(v[lo:hi]),
                                                     (hi - lo) * d, lr, b1, b2, eps, ptr(step), self.n_others,
                                                     self._tables["table"].others(off), stream_ptr()), "adam_step_peer")
        self.handshake()
