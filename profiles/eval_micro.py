"""one full-rank evaluation of LightGCN on the C2 shape through trainer.eval('test') (for launch lists / timing)"""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "inductive-recommendation_b200"), REPO]
import dataset as D  # noqa: E402
import model as M  # noqa: E402
import trainer as T  # noqa: E402
from b200rec import synth  # noqa: E402

dev = "cuda"
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
g = synth.generate_named(wl, device=dev)
ds = D.get_dataset({"name": "SyntheticDataset", "device": dev, "graph": g})
_, _, _, d, L = synth.SHAPES[wl]
m = M.get_model({"name": "LightGCN", "embedding_size": d, "n_layers": L, "device": dev}, ds)
tr = T.get_trainer({"name": "BPRTrainer", "optimizer": "Adam", "lr": 1e-3, "l2_reg": 1e-4, "device": dev, "n_epochs": 1,
                    "batch_size": 2048, "dataloader_num_workers": 0, "test_batch_size": 512, "topks": [1, 5, 10, 15, 20]}, ds, m)
for i in range(3):
    m._rep_cache = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.eval("test")
    torch.cuda.synchronize()
    print("eval %d: %.2f ms  (%.2f M users/s)" % (i, (time.perf_counter() - t0) * 1e3, ds.n_users / (time.perf_counter() - t0) / 1e6))
if os.environ.get("PROFILE", "0") == "1":  # host-side cost of one evaluation (cumulative, python level)
    import cProfile
    import pstats
    m._rep_cache = None
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    tr.eval("test")
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
