"""C4-shape sweep of the L2 residency hints: layer time against the size of the hot set kept in the persisting L2 set-aside.
usage: python profiles/spmm_hot_sweep.py [c4] [dims=128,16] [hot_mbs=0,32,48,64,80,96] [iters=5]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "inductive-recommendation_b200"), REPO]
from b200rec import graph, ops, synth  # noqa: E402

arg = lambda i, dflt: sys.argv[i] if len(sys.argv) > i else dflt  # noqa: E731
wl = arg(1, "c4")
dims = [int(x) for x in arg(2, "128,16").split(",")]
hot_mbs = [float(x) for x in arg(3, "0,32,48,64,80,96").split(",")]
iters = int(arg(4, "5"))
g = synth.generate_named(wl, device="cuda", heldout=False)
rows = torch.repeat_interleave(torch.arange(g.n_users, device="cuda"), g.train_indptr[1:] - g.train_indptr[:-1])
n = g.n_users + g.n_items
os.environ["B200REC_BLOCK_MB"] = "0"
op = graph.build_norm_adj(g.n_users, g.n_items, rows, g.train_items, "cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for d in dims:
    x = torch.randn((n, d), device="cuda")
    ref = None
    for cold in ("normal", "first"):
        for mb in hot_mbs:
            if mb == 0 and cold == "first":
                continue
            os.environ["B200REC_HOT_MB"], os.environ["B200REC_HOT_COLD"] = str(mb), cold
            y, acc = torch.empty_like(x), torch.zeros_like(x)
            h = op.hinted_for(d)
            ops.spmm(op, x, y=y, addend=acc, out=acc)
            torch.cuda.synchronize()
            if ref is None:
                ref = y.clone()
            else:
                assert torch.equal(ref, y), "hinted result differs"
            a, b = ev(), ev()
            a.record()
            for _ in range(iters):
                ops.spmm(op, x, y=y, addend=acc, out=acc)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            print("D=%3d hot=%5.1f MB cold=%-6s hot rows=%8d persisting=%6.1f MB : %8.3f ms/layer  gathered %6.0f GB/s" % (
                d, mb, cold, 0 if h is None else h.hot_rows, 0 if h is None else h.persist_bytes / 2 ** 20, ms,
                op.nnz * d * 4 / ms / 1e6), flush=True)
            del y, acc
    op._blocked.clear()
    del x, ref
    torch.cuda.empty_cache()
