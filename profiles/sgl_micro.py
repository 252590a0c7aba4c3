"""where the SGL step's time goes on the C2 shape: InfoNCE call, one view's propagation, whole step (L2-warm)"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "inductive-recommendation_b200"), REPO]
import dataset as D  # noqa: E402
import model as M  # noqa: E402
import trainer as T  # noqa: E402
from b200rec import ops, synth  # noqa: E402

dev = "cuda"
name = sys.argv[1] if len(sys.argv) > 1 else "SGL"
g = synth.generate_named("c2", device=dev)
ds = D.get_dataset({"name": "SyntheticDataset", "device": dev, "graph": g})
m = M.get_model({"name": name, "embedding_size": 64, "n_layers": 3, "aug_rate": 0.8, "device": dev}, ds)
tr = T.get_trainer({"name": name + "Trainer", "contrastive_reg": 0.1, "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4,
                    "device": dev, "n_epochs": 1, "batch_size": 2048, "dataloader_num_workers": 0, "test_batch_size": 512,
                    "topks": [1, 5, 10, 15, 20]}, ds, m)
m.train()
eng = tr._engine()


def timeit(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


print("step (graph): %.1f us" % timeit(eng.step))
q, k = eng.view_rep[0], eng.rep
print("infonce (B=2048, D=64, table form): %.1f us" % timeit(lambda: ops.infonce_fwd_bwd(
    q, k, eng.loss, eng.view_g[0], eng.g_rep, eng.nce_ws, rows=eng.batch, row_stride=3, n=2048)))
print("view fwd (3 layers, last restricted to users): %.1f us" % timeit(lambda: eng._propagate_fwd(
    eng.table, adj=m.norm_aug_adj1, flags=eng.user_flags, rep=eng.view_rep[0])))
print("view bwd: %.1f us" % timeit(lambda: eng._propagate_bwd(eng.view_grad, adj=m.norm_aug_adj1, flags=eng.user_flags,
                                                               g=eng.view_g[0])))
print("main fwd: %.1f us   main bwd: %.1f us" % (timeit(lambda: eng._propagate_fwd(eng.table)),
                                                 timeit(lambda: eng._propagate_bwd(eng.grad))))
print("grad.add_: %.1f us  zero 17.8MB: %.1f us  index_fill: %.1f us" % (
    timeit(lambda: eng.grad.add_(eng.view_grad)), timeit(lambda: eng.g_rep.zero_()),
    timeit(lambda: eng.user_flags.index_fill_(0, eng.batch[:, 0], 1))))
