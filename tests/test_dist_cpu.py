"""CPU: host-side logic of the multi-GPU decompositions (b200rec/dist.py) -- partition bounds, work-item slicing, and a
world_size-2 gloo run of the collectives the engine uses (block exchange, column gather, user-row gather, dot-product
all-reduce).  No kernels run here; the GPU tests compare the multi-rank results with the single-rank ones."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, REPO


def test_nnz_balanced_bounds():
    from b200rec.dist import nnz_balanced_bounds
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 50, 1000)
    lens[17] = 5000
    rp = np.zeros(1001, dtype=np.int64)
    np.cumsum(lens, out=rp[1:])
    for world in (1, 2, 4, 8):
        b = nnz_balanced_bounds(rp, world)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == world + 1 and all(x <= y for x, y in zip(b, b[1:]))
        per = [rp[b[p + 1]] - rp[b[p]] for p in range(world)]
        assert sum(per) == rp[-1]
        assert max(per) <= rp[-1] / world + 5000 + 50   # within one (hub) row of the ideal share
    assert nnz_balanced_bounds(np.zeros(5, dtype=np.int64), 4) == [0, 0, 0, 0, 4]  # empty graph


def test_dim_shard_ranges():
    from b200rec.dist import DimShard
    for world in (1, 2, 4, 8):
        cols = [DimShard(r, world, min_cols=1).cols(64) for r in range(world)]
        assert cols[0][0] == 0 and cols[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
        ur = [DimShard(r, world).user_range(1001) for r in range(world)]
        assert ur[0][0] == 0 and ur[-1][1] == 1001 and all(a[1] == b[0] for a, b in zip(ur, ur[1:]))
    with pytest.raises(AssertionError):
        DimShard(0, 3).cols(64)
    # the column split stops at min_cols: 8 ranks, D=64, min 16 columns -> 4-way shards x 2 replicas (no process group
    # is needed to compute the coordinates when the split covers the whole world)
    s = DimShard(5, 8, min_cols=8)
    s.configure(64)
    assert (s.world, s.rank) == (8, 5)
    # default narrowest shard: 32 columns (128-byte rows) for either kind of table; what is left of the world becomes
    # replicas -- identical ones for an L2-resident table, row-partitioned ones (row_partition) for one that streams from HBM
    small = DimShard(1, 2).configure(64, n_rows=69716)             # C2: 17.8 MB table
    assert (small.min_cols, small.world, small.rank, small.n_replicas) == (32, 2, 1, 1)
    big = DimShard(3, 4).configure(128, n_rows=3_000_000)          # C4 on 4 GPUs: 4 x 32 columns, no replicas
    assert (big.min_cols, big.world, big.rank, big.n_replicas) == (32, 4, 3, 1) and big.row_partition(None, 32) is None


def _worker(rank, world, port, out):
    sys.path[:0] = [PKG, REPO]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200rec.dist import DimShard, RowPartition, nnz_balanced_bounds
    shard = DimShard(rank, world, min_cols=1).configure(8)
    # column gather: every rank ends with the full table
    full = torch.arange(6 * 8, dtype=torch.float32).reshape(6, 8)
    lo, hi = shard.cols(8)
    assert torch.equal(shard.gather_cols(full[:, lo:hi].contiguous()), full)
    # dot-product all-reduce (the only coupling of a dimension-sharded BPR step)
    u, v = torch.randn(5, 8, generator=torch.Generator().manual_seed(1)), torch.randn(5, 8, generator=torch.Generator().manual_seed(2))
    part = (u[:, lo:hi] * v[:, lo:hi]).sum(1)
    assert torch.allclose(shard.all_reduce_sum(part.clone()), (u * v).sum(1), atol=1e-6)
    # user-sharded evaluation rows
    ids = torch.arange(11 * 3, dtype=torch.int32).reshape(11, 3)
    a, b = shard.user_range(11)
    assert torch.equal(shard.gather_user_rows(ids[a:b].contiguous(), 11), ids)
    # row-block exchange with uneven blocks
    class _Adj:  # only what RowPartition reads
        rowptr = torch.tensor([0, 1, 1, 9, 10, 12, 12, 20], dtype=torch.int32)
        n_rows, phase_split = 7, 3
        def row_slice(self, ranges):
            return ranges
    for split in (0, 3):  # one contiguous block per rank / one user block + one item block per rank
        _Adj.phase_split = split
        part = RowPartition(_Adj(), rank, world)
        covered = sorted(r for blk in part.blocks for lo, hi in blk for r in range(lo, hi))
        assert covered == list(range(7)) and len(part.blocks[rank]) == (2 if split else 1)
        table = torch.arange(7 * 4, dtype=torch.float32).reshape(7, 4)
        mine = torch.full_like(table, -1.0)
        for lo, hi in part.my_blocks:
            mine[lo:hi] = table[lo:hi]
        part.exchange(mine)
        assert torch.equal(mine, table)
    # replicas: with min_cols = D the columns are not split, each rank is its own shard group
    rep = DimShard(rank, world, min_cols=8).configure(8)
    assert (rep.world, rep.rank, rep.replica) == (1, 0, rank) and rep.cols(8) == (0, 8)
    assert torch.equal(rep.gather_cols(full), full)
    assert torch.equal(rep.gather_user_rows(ids[a:b].contiguous(), 11), ids)   # users still shard over the world
    out.put(rank)
    dist.destroy_process_group()


def test_world_size_2_gloo_collectives():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]


def _worker_replicas(rank, world, port, out):
    """4 ranks, D = 8, narrowest shard 4 columns -> two 2-way shard groups (replicas): the bench layout on 4 / 8 GPUs"""
    sys.path[:0] = [PKG, REPO]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200rec.dist import DimShard
    s = DimShard(rank, world, min_cols=4).configure(8)
    assert (s.world, s.rank, s.replica) == (2, rank % 2, rank // 2) and s.cols(8) == ((rank % 2) * 4, (rank % 2) * 4 + 4)
    # the per-step all-reduce stays inside the replica's shard group ...
    t = s.all_reduce_sum(torch.tensor([float(rank)]))
    assert float(t) == float(4 * (rank // 2) + 1)          # ranks {0,1} -> 1, ranks {2,3} -> 5
    # ... the column gather too, while evaluation rows are sharded over the whole world
    full = torch.arange(3 * 8, dtype=torch.float32).reshape(3, 8)
    lo, hi = s.cols(8)
    assert torch.equal(s.gather_cols(full[:, lo:hi].contiguous()), full)
    ids = torch.arange(10 * 2, dtype=torch.int32).reshape(10, 2)
    a, b = s.user_range(10)
    assert (a, b) == (min(10, rank * 3), min(10, rank * 3 + 3))
    assert torch.equal(s.gather_user_rows(ids[a:b].contiguous(), 10), ids)
    # hybrid layout: the ranks that hold the same columns in the two replicas ({0,2} and {1,3}) split the rows
    assert s.n_replicas == 2 and s.peer_group is not None
    t = torch.tensor([float(rank)])
    dist.all_reduce(t, group=s.peer_group)
    assert float(t) == float(2 * (rank % 2) + 2)           # {0,2} -> 2, {1,3} -> 4

    class _Adj:  # only what RowPartition reads
        rowptr = torch.tensor([0, 1, 1, 9, 10, 12, 12, 20], dtype=torch.int32)
        n_rows, phase_split = 7, 3

        def row_slice(self, ranges):
            return ranges
    part = s.row_partition(_Adj(), 4, kind="row")          # NCCL/gloo block exchange among the peers
    assert part.world == 2 and part.rank == rank // 2
    table = torch.arange(7 * 4, dtype=torch.float32).reshape(7, 4) + 100 * (rank % 2)   # each column shard its own values
    mine = torch.full_like(table, -1.0)
    for lo2, hi2 in part.my_blocks:
        mine[lo2:hi2] = table[lo2:hi2]
    part.exchange(mine)
    assert torch.equal(mine, table)
    out.put(rank)
    dist.destroy_process_group()


def test_world_size_4_gloo_replicated_shard_groups():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_replicas, args=(r, 4, port, out)) for r in range(4)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(4)) == [0, 1, 2, 3]
