#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json: BPR epochs/s, SpMM HBM GB/s, eval users/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl reference]

A "step" is one BPR training iteration (batch 2048: sample -> L x SpMM + layer mean -> BPR loss/grad -> L x SpMM
backward -> Adam) on a synthetic graph of a BASELINE.json shape; default workload c2 (Yelp2018-shaped, LightGCN,
3 layers, dim 64).  `value` = epochs/s = 1 / (steps_per_epoch * s_per_step) with all inputs resident in HBM
(device sampler).  `e2e` = the same metric through the trainer-facing engine call with HOST batches: every step
copies a pinned int64 [B,3] batch host->device and reads the loss back.  One JSON line on stdout (rank 0).

`--impl reference` times the oracle port (oracle/ref_port.py: the reference's algorithm restated on torch-CPU with
an MKL CSR SpMM, all host threads) on the same workload; the reference itself is pure Python + DGL and cannot be
installed or shipped to the GPU box (no DGL wheel, /root/reference is not present there).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "inductive-recommendation_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 2048
LR, L2_REG = 1e-3, 1e-4  # config.py:13-14 (LightGCN + BPRTrainer)
TOPKS = [1, 5, 10, 15, 20]
WORKLOAD_DOC = {
    "c1": "LightGCN L=3 D=64, synthetic Gowalla-shaped graph 29858x40981, 1027370 train edges",
    "c2": "LightGCN L=3 D=64, synthetic Yelp2018-shaped graph 31668x38048, 1561406 train edges",
    "c3": "LightGCN L=3 D=64, synthetic Amazon-book-shaped graph 52643x91599, 2984108 train edges",
    "c4": "LightGCN L=4 D=128, synthetic power-law graph 2000000x1000000, 100000000 train edges",
}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_graph(workload, device):
    from b200rec import synth
    return synth.generate_named(workload, seed=0, device=device)


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_port_step_time(graph, workload, n_steps, warmup, seed=2021):
    """oracle port: LightGCN train step on torch-CPU (CSR/MKL), all host threads.  Returns (s_per_step, threads)."""
    from b200rec import synth
    from oracle import oracle_c as oc
    from oracle import ref_port as rp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    _, _, _, d, n_layers = synth.SHAPES[workload]
    ptr = graph.train_indptr.cpu().numpy()
    items = graph.train_items.cpu().numpy()
    users, items64 = rp.pairs_from_csr(ptr, items)
    torch.manual_seed(seed)
    emb0 = (0.1 * torch.randn(graph.n_users + graph.n_items, d)).numpy()
    port = rp.LightGCNPort(graph.n_users, graph.n_items, users, items64, emb0, n_layers).train()
    opt = torch.optim.Adam(port.parameters(), lr=LR)
    p32, i32 = ptr.astype(np.int32), items.astype(np.int32)
    batches = [torch.from_numpy(oc.bpr_sample(p32, i32, graph.n_users, graph.n_items, seed, s, BATCH))
               for s in range(n_steps + warmup)]
    for b in batches[:warmup]:
        rp.train_step(port, opt, b, L2_REG)
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        rp.train_step(port, opt, b, L2_REG)
    return (time.perf_counter() - t0) / max(n_steps, 1), threads, port


def cpu_port_eval_rate(graph, port, n_users=1024):
    """full-rank evaluation of the first `n_users` users by the oracle port, the way the reference does it (trainer.py:
    146-170: every 512-user batch re-propagates, dense scores, Python exclusion lists, topk): users/s on the host cores"""
    from oracle import ref_port as rp
    n = min(n_users, graph.n_users)
    train, val, test = graph.lists("train")[:n], graph.lists("val")[:n], graph.lists("test")[:n]
    t0 = time.perf_counter()
    rp.evaluate(port, train, val, test, "test", TOPKS, test_batch_size=512, n_users=n)
    return n / (time.perf_counter() - t0), n


def run_reference(args, rank):
    if rank != 0:
        return
    workload = args.workload or "c2"
    graph = build_graph(workload, "cpu")
    n_train = int(graph.train_items.numel())
    steps_per_epoch = (n_train + BATCH - 1) // BATCH
    s_per_step, threads, port = cpu_port_step_time(graph, workload, args.steps, args.warmup)
    value = 1.0 / (steps_per_epoch * s_per_step)
    eval_rate, eval_n = cpu_port_eval_rate(graph, port)
    sample = "%d timed train steps (batch %d) of %d per epoch, extrapolated" % (args.steps, BATCH, steps_per_epoch)
    line = {"impl": "reference", "metric": "bpr_epochs_per_sec", "value": value, "unit": "epochs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload + ": " + WORKLOAD_DOC[workload], "batch": BATCH,
                       "steps_per_epoch": steps_per_epoch},
            "cpu_baseline": {"value": value, "unit": "epochs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "eval": {"users_per_s_e2e": eval_rate, "sample": "first %d users, test split, K=%d, 512-user batches" % (eval_n, max(TOPKS))}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- own arm
def run_own(args, rank, world):
    import dataset as D
    import model as M
    import trainer as T
    from b200rec import _abi, ops, synth

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    workload = args.workload or "c2"
    _, _, _, d, n_layers = synth.SHAPES[workload]
    graph = build_graph(workload, dev)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": dev, "graph": graph})
    torch.manual_seed(2021)  # same seed on every rank: the shards are slices of one initialisation
    partition = None
    model_cfg = {"lightgcn": {"name": "LightGCN", "embedding_size": d, "n_layers": n_layers},
                 "igcn": {"name": "IGCN", "embedding_size": d, "n_layers": n_layers, "dropout": 0.3, "feature_ratio": 1.0},
                 "mf": {"name": "MF", "embedding_size": d},
                 "sgl": {"name": "SGL", "embedding_size": d, "n_layers": n_layers, "aug_rate": 0.8},
                 "half": {"name": "HALF", "embedding_size": d, "n_layers": n_layers, "aug_rate": 0.8}}[args.model]
    m = M.get_model(dict(model_cfg, device=dev), ds)
    if world > 1:
        from b200rec.dist import DimShard, PeerRowPartition, RowPartition, shard_model_dims
        if args.parallelism == "row":
            partition = RowPartition(m.norm_adj, rank, world)
        elif args.parallelism == "peer":
            partition = PeerRowPartition(m.norm_adj, rank, world, d)
        else:
            shard_model_dims(m, DimShard(rank, world))
            d = m.embedding_size
    trainer_name = {"igcn": "IGCNTrainer", "sgl": "SGLTrainer", "half": "HALFTrainer"}.get(args.model, "BPRTrainer")
    tr = T.get_trainer({"name": trainer_name, "contrastive_reg": 0.1, "optimizer": "Adam", "lr": LR,
                        "l2_reg": 0.0 if args.model == "igcn" else L2_REG, "aux_reg": 0.01, "device": dev,
                        "n_epochs": 1, "batch_size": BATCH, "dataloader_num_workers": 0, "test_batch_size": 512,
                        "topks": TOPKS, "partition": partition}, ds, m)
    m.train()
    eng = tr._engine()
    steps_per_epoch = tr.steps_per_epoch()
    K, W = args.steps, max(args.warmup, 3)
    l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
    flush_buf = torch.empty(max(2 * l2_bytes, 256 << 20), dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_steps(fn, n, flush):
        """device time of n calls of fn, CUDA events on the launching stream; L2 flushed (untimed) between calls"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            if flush:
                flush_buf.fill_(1)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in ev]

    for _ in range(W):
        eng.step()
    barrier()
    launches0 = _abi.launch_count()
    graph_replays0 = eng.steps_done
    with ClockSampler(local) as clocks:
        times = timed_steps(eng.step, K, flush=True)
        barrier()
        # the same K steps back to back (L2 warm: the steady state of a real epoch), one event pair
        t_warm = timed_steps(lambda: [eng.step() for _ in range(K)], 1, flush=False)[0] / K
    total_ms = float(sum(times))
    if world > 1:
        t = torch.tensor([total_ms, t_warm], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms, t_warm = float(t[0]), float(t[1])
    ms_per_step = total_ms / K
    value = 1e3 / (ms_per_step * steps_per_epoch)
    kernels_per_step = eng.kernels_per_step()
    gpu_launches = kernels_per_step * K
    assert _abi.launch_count() >= launches0  # replays do not pass through the host counter

    # ---- e2e: host batches through the engine's public step(), H2D of the batch + D2H of the loss every step ----
    host_batches = []
    st = torch.zeros(1, dtype=torch.int64, device=dev)
    for s in range(K + W):
        st.fill_(10_000_000 + s)
        host_batches.append(ops.bpr_sample(eng.user_ptr, eng.user_items, ds.n_users, ds.n_items, 2021, st, BATCH)
                            .cpu().pin_memory())
    for b in host_batches[:W]:
        eng.step(host_batch=b)
        eng.last_loss()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t0.record()
    for b in host_batches[W:]:
        flush_buf.fill_(1)
        eng.step(host_batch=b)
        loss = eng.last_loss()  # device->host read of the step's loss (the reference's loss.item(), trainer.py:428)
    t1.record()
    torch.cuda.synchronize()
    e2e_ms = max(t0.elapsed_time(t1), 1e3 * (time.perf_counter() - wall0)) / K
    # the flush is inside this region (it cannot be hoisted out of a host-synchronous loop): subtract its measured cost
    flush_ms = float(np.median(timed_steps(lambda: flush_buf.fill_(1), 10, flush=False)))
    e2e_ms_net = max(e2e_ms - flush_ms, 1e-6)
    if world > 1:
        t = torch.tensor([e2e_ms_net], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms_net = float(t[0])
    e2e_value = 1e3 / (e2e_ms_net * steps_per_epoch)

    # ---- roofline of the dominant kernel: one SpMM layer (propagation forward layer k: Y = A X, acc += Y) ----
    n = ds.n_users + ds.n_items
    adj = m.norm_adj if partition is None else partition.local_op
    nnz_local = int(adj.item_end[:adj.n_items].sum() - adj.item_start[:adj.n_items].sum()) if partition is not None else adj.nnz
    rows_local = n if partition is None else partition.n_local_rows
    x = torch.randn((n, d), device=dev)
    y = torch.empty_like(x)
    acc = torch.zeros_like(x)
    spmm_call = lambda: ops.spmm(adj, x, y=y, addend=acc, out=acc)  # noqa: E731
    for _ in range(3):
        spmm_call()
    spmm_ms = float(np.median(timed_steps(spmm_call, 30, flush=True)))
    spmm_ms_warm = timed_steps(lambda: [spmm_call() for _ in range(30)], 1, flush=False)[0] / 30
    # algorithmic bytes per launch: nnz*(4+4) + (N+1)*4 + 2*N*D*4 ; + N*D*4 for the fused layer-sum read+write/2
    bytes_min = nnz_local * 8 + (rows_local + 1) * 4 + (n + rows_local) * d * 4
    bytes_gather = nnz_local * 8 + nnz_local * d * 4 + rows_local * d * 4
    peak, peak_src = peaks()
    achieved = bytes_min / (spmm_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "spmm_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(workload)
    roofline = {"bound": "hbm", "kernel": "spmm_items_kernel (one propagation layer, D=%d)" % d, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_min": bytes_min, "launch_ms": spmm_ms, "launch_ms_l2_warm": spmm_ms_warm,
                "effective_gather_gbs": bytes_gather / (spmm_ms * 1e-3) / 1e9,
                "effective_gather_gbs_l2_warm": bytes_gather / (spmm_ms_warm * 1e-3) / 1e9}

    # ---- full-rank evaluation (users/s): representation + fused score/mask/top-20 + metrics ----
    eval_info = None
    if workload != "c4":  # every rank takes part: the sweep is user-sharded and gathers the columns collectively
        m.eval()
        tr.eval("test")
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        _, metrics, _ = tr.eval("test")
        torch.cuda.synchronize()
        eval_s = time.perf_counter() - w0
        ke = timed_steps(lambda: tr.recommend_all("test"), 3, flush=True)
        d_full = getattr(m, "full_embedding_size", m.embedding_size)
        score_flops = 2.0 * ds.n_users * ds.n_items * d_full
        pk = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
        tf_peak = float(pk.get("bf16_tflops", 1590.0))
        tf = score_flops / (min(ke) * 1e-3) / 1e12 * (1.0 / world)  # per GPU: the sweep is user-sharded
        eval_info = {"users_per_s_e2e": ds.n_users / eval_s, "users_per_s_kernels": ds.n_users / (min(ke) * 1e-3),
                     "k": max(TOPKS), "recall@20": float(metrics["Recall"][20]), "ndcg@20": float(metrics["NDCG"][20]),
                     "precision": tr.eval_precision,
                     "score_roofline": {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s",
                                        "frac": tf / tf_peak, "note": "whole recommend_all sweep (propagation + prep + "
                                        "tcgen05 score/candidates + exact re-score) per GPU; epilogue-latency-bound, see DESIGN 4.6"},
                     "includes": "trainer.eval('test'): get_rep + score/mask/top-K + fused rank-metrics pass + D2H of the metric sums"}

    # ---- CPU baseline beside it (rank 0, N=1): bounded sample of the same workload on the host cores ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "lightgcn":
        n_cpu = 3 if workload != "c4" else 1
        s_cpu, threads, _ = cpu_port_step_time(graph, workload, n_cpu, 1)
        cpu_baseline = {"value": 1.0 / (steps_per_epoch * s_cpu), "unit": "epochs/s", "cores": threads, "kind": "port",
                        "sample": "%d timed train steps (batch %d) of %d per epoch on torch-CPU CSR/MKL, extrapolated"
                                  % (n_cpu, BATCH, steps_per_epoch), "ms_per_step": 1e3 * s_cpu}
    if rank == 0:
        line = {"metric": "bpr_epochs_per_sec", "value": value, "unit": "epochs/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload + ": " + WORKLOAD_DOC[workload].replace("LightGCN", model_cfg["name"]), "batch": BATCH,
                           "steps_per_epoch": steps_per_epoch, "optimizer": "Adam", "sampler": "device (Philox)",
                           "l2": "flushed between timed steps (write of %d MiB)" % (flush_buf.numel() >> 20),
                           "parallelism": "single GPU" if world == 1 else (
                               ("row-partitioned graph x%d, exchange fused into the SpMM / Adam epilogues (NVLink peer stores) "
                                "+ signal/wait hand-shake" % world) if args.parallelism == "peer" else
                               "row-partitioned graph x%d, per-layer NCCL block exchange" % world if partition is not None else
                               "embedding dimension sharded x%d (%d columns per GPU) x %d replicas, one [B,3] all-reduce per "
                               "step; evaluation user-sharded x%d" % (m._dim_shard.world, d, world // m._dim_shard.world, world))},
                "ms_per_step_l2_warm": t_warm, "epochs_per_sec_l2_warm": 1e3 / (t_warm * steps_per_epoch),
                "e2e": {"value": e2e_value, "unit": "epochs/s", "h2d_bytes_per_step": BATCH * 3 * 8,
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms_net, "last_loss": loss},
                "gpu_launches": gpu_launches, "kernels_per_step": kernels_per_step,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "eval": eval_info, "clocks": clocks.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        eng.close()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c1", "c2", "c3", "c4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="lightgcn", choices=["lightgcn", "igcn", "mf", "sgl", "half"],
                    help="lightgcn is the headline; igcn = inductive template-feature layer + propagation (config.py:18-23)")
    ap.add_argument("--parallelism", default="dim", choices=["dim", "row", "peer"],
                    help="multi-GPU decomposition: embedding-dimension sharding (default), the row partition with NCCL "
                         "block exchange, or the row partition with the exchange fused into the kernels over peer memory")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU port)")
    run_own(args, rank, world)


if __name__ == "__main__":
    main()
