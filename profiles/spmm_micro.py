"""micro-benchmark of one propagation layer (Y = A X, acc += Y) at several embedding widths; used for ncu captures.
usage: python profiles/spmm_micro.py c2 8,16,32,64 [iters]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "inductive-recommendation_b200"), REPO]
from b200rec import graph, ops, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dims = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "64").split(",")]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
g = synth.generate_named(wl, device="cuda")
rows = torch.repeat_interleave(torch.arange(g.n_users, device="cuda"), g.train_indptr[1:] - g.train_indptr[:-1])
n = g.n_users + g.n_items
for d in dims:
    op = graph.build_norm_adj(g.n_users, g.n_items, rows, g.train_items, "cuda", d=d)
    x = torch.randn((n, d), device="cuda")
    y = torch.empty_like(x)
    acc = torch.zeros_like(x)
    for _ in range(3):
        ops.spmm(op, x, y=y, addend=acc, out=acc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        ops.spmm(op, x, y=y, addend=acc, out=acc)
    b.record()
    torch.cuda.synchronize()
    print("%s D=%d items=%d long=%d chunk=%d: %.1f us/layer (L2-warm)" % (wl, d, op.n_items, op.n_long, op.chunk,
                                                                         a.elapsed_time(b) / iters * 1e3), flush=True)
    if os.environ.get("MASKED", "0") == "1":
        # the two restricted layers of a BPR step: flags from one device-sampled batch of 2048 triples
        ptr32, idx32 = g.train_indptr.to(torch.int32), g.train_items.to(torch.int32)
        step = torch.zeros(1, dtype=torch.int64, device="cuda")
        batch = ops.bpr_sample(ptr32, idx32, g.n_users, g.n_items, 2021, step, 2048)
        flags = torch.zeros(n, dtype=torch.uint8, device="cuda")
        ops.mark_rows(batch, g.n_users, flags)
        gs = torch.zeros_like(x)
        gs[flags.bool()] = torch.randn((int(flags.sum()), d), device="cuda")
        frac = float((flags[op.colidx.long() & 0x7fffffff] != 0).float().mean())
        for name, kw, src in (("dst-masked", dict(dst_flags=flags), x), ("src-masked", dict(src_flags=flags), gs)):
            for _ in range(3):
                ops.spmm(op, src, y=y, addend=acc, out=acc, **kw)
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                ops.spmm(op, src, y=y, addend=acc, out=acc, **kw)
            b.record()
            torch.cuda.synchronize()
            print("   %s: %.1f us/layer  (flagged rows %d, edges with a flagged source %.1f %%)" % (
                name, a.elapsed_time(b) / iters * 1e3, int(flags.sum()), 100 * frac), flush=True)
