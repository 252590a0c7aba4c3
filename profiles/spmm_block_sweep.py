"""C4-shape sweep of the column-blocked SpMM: layer time against block size and sweep width, per table width.
usage: python profiles/spmm_block_sweep.py [c4] [dims=128,16] [block_mbs=0,32,48,64,96] [sweeps=128,64,32] [iters=5]
Prints one line per configuration; the table is 3M x D fp32 (D = 128: 1.5 GB), every configuration gives identical bits
(asserted against the first one)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "inductive-recommendation_b200"), REPO]
from b200rec import graph, ops, synth  # noqa: E402

arg = lambda i, dflt: sys.argv[i] if len(sys.argv) > i else dflt  # noqa: E731
wl = arg(1, "c4")
dims = [int(x) for x in arg(2, "128,16").split(",")]
block_mbs = [float(x) for x in arg(3, "0,32,48,64,96").split(",")]
sweeps = [int(x) for x in arg(4, "128,64,32").split(",")]
iters = int(arg(5, "5"))
g = synth.generate_named(wl, device="cuda", heldout=False)
rows = torch.repeat_interleave(torch.arange(g.n_users, device="cuda"), g.train_indptr[1:] - g.train_indptr[:-1])
n = g.n_users + g.n_items
os.environ["B200REC_BLOCK_MB"] = "0"
op = graph.build_norm_adj(g.n_users, g.n_items, rows, g.train_items, "cuda")
print("%s: n=%d nnz=%d items=%d long=%d chunk=%d" % (wl, n, op.nnz, op.n_items, op.n_long, op.chunk), flush=True)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for d in dims:
    x = torch.randn((n, d), device="cuda")
    ref = None
    for mb in block_mbs:
        for sw in ([d] if mb == 0 else sorted({min(s, d) for s in sweeps}, reverse=True)):
            os.environ["B200REC_BLOCK_MB"], os.environ["B200REC_SWEEP_D"] = str(mb), str(sw)
            y = torch.empty_like(x)
            acc = torch.zeros_like(x)
            blk = op.blocked_for(d)
            ops.spmm(op, x, y=y, addend=acc, out=acc)
            torch.cuda.synchronize()
            if ref is None:
                ref = y.clone()
            else:
                assert torch.equal(ref, y), "blocked result differs"
            a, b = ev(), ev()
            a.record()
            for _ in range(iters):
                ops.spmm(op, x, y=y, addend=acc, out=acc)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            bytes_min = op.nnz * 8 + (n + 1) * 4 + 2 * n * d * 4
            print("D=%3d block=%5.1f MB sweep=%3d passes=%3d records=%10d window=%4d : %8.3f ms/layer  bytes_min/t = %6.0f GB/s  "
                  "gathered %6.0f GB/s" % (
                      d, mb, sw, 0 if blk is None else blk[0].n_passes * (d // blk[1]), op.nnz if blk is None else blk[0].n_records,
                      0 if blk is None else blk[0].window, ms, bytes_min / ms / 1e6, op.nnz * d * 4 / ms / 1e6), flush=True)
            if blk is not None:
                op._blocked.clear()  # free the plan before building the next one
            del y, acc
    del x, ref
    op._carry.clear()
    torch.cuda.empty_cache()
