"""micro-benchmark of full-rank score + mask + top-K: exact fp32 CUDA-core path vs the tcgen05 path.
usage: python profiles/score_micro.py n_users n_items d k"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "inductive-recommendation_b200"), REPO]
from b200rec import ops  # noqa: E402

nu, ni, d, k = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (31668, 38048, 64, 20)
g = torch.Generator(device="cuda").manual_seed(0)
rep_u = 0.1 * torch.randn((nu, d), device="cuda", generator=g)
rep_i = 0.1 * torch.randn((ni, d), device="cuda", generator=g)
users = torch.arange(nu, device="cuda")
tc_only = os.environ.get("TC_ONLY", "0") == "1"  # for ncu captures: the tcgen05 path only, one chunk size
for prec in ((1,) if tc_only else (0, 1)):
    for chunk in ((65536,) if tc_only else (16384, 65536)):
        def run():
            for s in range(0, nu, chunk):
                ops.score_topk(rep_u, users[s:s + chunk].contiguous(), rep_i, k, precision=prec)
        run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print("precision %d chunk %6d: %8.3f ms  %10.0f users/s  %7.1f TFLOP/s  overflow rows %s" % (
            prec, chunk, ms, nu / ms * 1e3, 2.0 * nu * ni * d / ms / 1e9, getattr(ops.score_topk, "last_overflow", "-")), flush=True)
