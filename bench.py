#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json: BPR epochs/s, SpMM HBM GB/s, eval users/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

A "step" is one BPR training iteration (batch 2048: sample -> L x SpMM + layer mean -> BPR loss/grad -> L x SpMM
backward -> Adam) on a synthetic graph of a BASELINE.json shape.  Default workload: c4, the largest configuration that
fits one GPU (2M x 1M power-law graph, 100M train edges, LightGCN L=4 D=128); its full-rank evaluation is BASELINE's
config 5 (2M x 1M score sweep + masked top-20).  `value` = epochs/s = 1 / (steps_per_epoch * s_per_step) with all
inputs resident in HBM (device sampler).  `e2e` = the same metric through the trainer-facing engine call with HOST
batches: every step copies a pinned int64 [B,3] batch host->device and reads the loss back.  One JSON line on stdout
(rank 0).  `--workload c5` times the evaluation sweep alone (metric eval_users_per_sec).

`--impl reference` times the oracle port (oracle/ref_port.py: the reference's algorithm restated on torch-CPU with
an MKL CSR SpMM, all host threads) on the same graph (the generator is device-independent); the reference itself is
pure Python + DGL and cannot be installed or shipped to the GPU box (no DGL wheel, /root/reference is not present
there).  On c4 one CPU step is ~8-30 s, so a reference "step" is a bounded sample of it (see CpuPort).
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
DEFAULT_WORKLOAD = "c4"
WORKLOAD_DOC = {
    "c1": "LightGCN L=3 D=64, synthetic Gowalla-shaped graph 29858x40981, 1027370 train edges",
    "c2": "LightGCN L=3 D=64, synthetic Yelp2018-shaped graph 31668x38048, 1561406 train edges",
    "c3": "LightGCN L=3 D=64, synthetic Amazon-book-shaped graph 52643x91599, 2984108 train edges",
    "c4": "LightGCN L=4 D=128, synthetic power-law graph 2000000x1000000, 100000000 train edges (BASELINE configs 4+5)",
    "c5": "full-rank evaluation sweep 2000000 users x 1000000 items, D=128, train+val masked top-20 (BASELINE config 5, graph of c4)",
}
BIG = ("c4", "c5")


_T0 = time.time()


def _log(msg):
    """progress on stderr (stdout carries the one JSON line)"""
    if int(os.environ.get("RANK", 0)) == 0:
        print("[bench %7.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


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
            time.sleep(0.1)

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


def graph_name(workload):
    return "c4" if workload == "c5" else workload


def build_graph(workload, device):
    """the same graph whichever device builds it (b200rec.synth is integer-only); both arms use the GPU when there is one"""
    from b200rec import synth
    return synth.generate_named(graph_name(workload), seed=0, device=device)


# ---------------------------------------------------------------------------------------------------- reference arm
class CpuPort:
    """oracle port of the LightGCN training step on torch-CPU (CSR/MKL), all host threads.

    Small workloads: a timed step is one full train step (trainer.py:412-429).  c4: a full step is 8 SpMM layers over
    200M entries (tens of seconds), so a timed step is a BOUNDED SAMPLE -- the complete train step of a ONE-layer port
    (1 forward + 1 backward SpMM over the whole graph, the batch gathers, loss, backward, dense Adam over the 3M x 128
    table), and the full step is estimated as  t_L1 + (L-1) * t_pair  with t_pair = one forward + one backward SpMM
    timed alone.  The estimate leaves out the (L+1)-way stack/mean of the deeper model, i.e. it favours the CPU."""

    def __init__(self, graph, workload, seed=2021):
        from b200rec import synth
        from oracle import oracle_c as oc
        from oracle import ref_port as rp
        self.rp, self.oc = rp, oc
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        _, _, _, d, n_layers = synth.SHAPES[graph_name(workload)]
        self.n_layers, self.sampled = n_layers, workload in BIG
        self.graph = graph
        self.ptr = graph.train_indptr.cpu().numpy()
        self.items = graph.train_items.cpu().numpy()
        torch.manual_seed(seed)
        emb0 = (0.1 * torch.randn(graph.n_users + graph.n_items, d)).numpy()
        adj = rp.norm_adjacency_from_csr(graph.n_users, graph.n_items, self.ptr, self.items)  # == norm_adjacency, no COO sort
        self.port = rp.LightGCNPort(graph.n_users, graph.n_items, None, None, emb0, 1 if self.sampled else n_layers,
                                    adj_sp=adj).train()
        self.opt = torch.optim.Adam(self.port.parameters(), lr=LR)
        self.p32, self.i32 = self.ptr.astype(np.int32), self.items.astype(np.int32)
        self.seed, self.step_no = seed, 0
        self.t_pair = None

    def _batch(self):
        g = self.graph
        b = torch.from_numpy(self.oc.bpr_sample(self.p32, self.i32, g.n_users, g.n_items, self.seed, self.step_no, BATCH))
        self.step_no += 1
        return b

    def measure_pair(self, reps=2):
        a = self.port.a
        x = self.port.embedding.weight.detach()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            y = torch.sparse.mm(a, x)
            torch.sparse.mm(a, y)
            ts.append(time.perf_counter() - t0)
        self.t_pair = float(np.median(ts))

    def step(self):
        """seconds of one (estimated) full train step"""
        b = self._batch()
        t0 = time.perf_counter()
        self.rp.train_step(self.port, self.opt, b, L2_REG)
        t = time.perf_counter() - t0
        if self.sampled:
            if self.t_pair is None:
                self.measure_pair()
            t += (self.n_layers - 1) * self.t_pair
        return t

    def sample_doc(self, n_steps, steps_per_epoch):
        if self.sampled:
            return ("%d timed samples; each = the complete train step of a 1-layer port (1 fwd + 1 bwd SpMM over all %d "
                    "entries, batch %d, dense Adam) + (L-1) x one fwd+bwd SpMM pair timed alone (%.2f s); extrapolated to "
                    "%d steps per epoch" % (n_steps, 2 * self.items.size, BATCH, self.t_pair or 0.0, steps_per_epoch))
        return "%d timed train steps (batch %d) of %d per epoch on torch-CPU CSR/MKL, extrapolated" % (n_steps, BATCH, steps_per_epoch)

    def eval_rate(self, n_users):
        """full-rank evaluation of the first `n_users` users the way the reference does it (trainer.py:146-170: every
        512-user batch re-propagates, dense scores, Python exclusion lists, topk): users/s on the host cores"""
        g = self.graph
        n = min(n_users, g.n_users)
        lists = lambda which: _head_lists(getattr(g, which + "_indptr"), getattr(g, which + "_items"), n)  # noqa: E731
        full_layers = self.port.n_layers
        self.port.n_layers = self.n_layers  # evaluation always runs the real depth
        t0 = time.perf_counter()
        self.rp.evaluate(self.port, lists("train"), lists("val"), lists("test"), "test", TOPKS, test_batch_size=512, n_users=n)
        dt = time.perf_counter() - t0
        self.port.n_layers = full_layers
        return n / dt, n


def _head_lists(indptr, items, n):
    ptr = indptr[: n + 1].cpu().numpy()
    idx = items[: int(ptr[-1])].cpu().numpy()
    return [idx[ptr[u]:ptr[u + 1]].tolist() for u in range(n)]


def run_reference(args, rank):
    if rank != 0:
        return
    workload = args.workload or DEFAULT_WORKLOAD
    dev = "cuda" if torch.cuda.is_available() else "cpu"  # input generation only; everything timed below is CPU work
    graph = build_graph(workload, dev)
    n_train = int(graph.train_items.numel())
    steps_per_epoch = (n_train + BATCH - 1) // BATCH
    _log("reference arm: graph ready, building the CPU port")
    port = CpuPort(graph, workload)
    _log("CPU port built (%d threads)" % port.threads)
    for _ in range(args.warmup):
        port.step()
    times = []
    for i in range(args.steps):
        times.append(port.step())
        _log("reference step %d: %.2f s" % (i, times[-1]))
    s_per_step = float(np.mean(times)) if times else float("nan")
    value = 1.0 / (steps_per_epoch * s_per_step)
    eval_rate, eval_n = port.eval_rate(256 if workload in BIG else 1024)
    _log("reference evaluation sample: %.1f users/s" % eval_rate)
    if workload == "c5":
        metric, unit, value_out = "eval_users_per_sec", "users/s", eval_rate
    else:
        metric, unit, value_out = "bpr_epochs_per_sec", "epochs/s", value
    line = {"impl": "reference", "metric": metric, "value": value_out, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload + ": " + WORKLOAD_DOC[workload], "batch": BATCH,
                       "steps_per_epoch": steps_per_epoch, "graph_fingerprint": "%016x" % _fingerprint(graph)},
            "cpu_baseline": {"value": value_out, "unit": unit, "cores": port.threads, "kind": "port",
                             "sample": port.sample_doc(args.steps, steps_per_epoch) if workload != "c5" else
                             "first %d users, test split, K=%d" % (eval_n, max(TOPKS))},
            "e2e": {"value": value_out, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "train": {"epochs_per_s": value, "ms_per_step": 1e3 * s_per_step},
            "eval": {"users_per_s_e2e": eval_rate, "sample": "first %d users, test split, K=%d, 512-user batches, re-propagation "
                     "per batch like trainer.py:150-170" % (eval_n, max(TOPKS))}}
    _emit(line)


def _fingerprint(graph):
    from b200rec import synth
    return synth.fingerprint(graph)


# ---------------------------------------------------------------------------------------------------- own arm
def small_shape_leg(dev, steps=20, warmup=5):
    """C2 (Yelp2018-shaped, L2-resident tables) beside the headline: the same step and the full-rank evaluation on the shape
    round 1 was quoted on, L2 flushed between timed steps, evaluation at the near-initial embeddings these few steps leave
    (the state in which round 1's global candidate band was slowest)."""
    import dataset as D
    import model as M
    import trainer as T
    from b200rec import synth
    _, _, _, d, n_layers = synth.SHAPES["c2"]
    graph = synth.generate_named("c2", seed=0, device=dev)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": dev, "graph": graph})
    torch.manual_seed(2021)
    m = M.get_model({"name": "LightGCN", "embedding_size": d, "n_layers": n_layers, "device": dev}, ds)
    tr = T.get_trainer({"name": "BPRTrainer", "optimizer": "Adam", "lr": LR, "l2_reg": L2_REG, "device": dev, "n_epochs": 1,
                        "batch_size": BATCH, "dataloader_num_workers": 0, "test_batch_size": 512, "topks": TOPKS}, ds, m)
    m.train()
    eng = tr._engine()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        eng.step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush_buf.fill_(1)
        a.record()
        eng.step()
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    m.eval()
    tr.eval("test")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        _, metrics, _ = tr.eval("test")
    torch.cuda.synchronize()
    eval_s = (time.perf_counter() - t0) / 3
    out = {"workload": "c2: " + WORKLOAD_DOC["c2"], "ms_per_step": ms, "epochs_per_s": 1e3 / (ms * tr.steps_per_epoch()),
           "kernels_per_step": eng.kernels_per_step(), "eval_users_per_s_e2e": ds.n_users / eval_s, "eval_ms": 1e3 * eval_s,
           "trained_steps": eng.steps_done, "recall@20": float(metrics["Recall"][20]),
           "l2": "flushed between timed steps (write of 256 MiB)"}
    eng.close()
    return out


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
    workload = args.workload or DEFAULT_WORKLOAD
    big = workload in BIG
    _, _, _, d, n_layers = synth.SHAPES[graph_name(workload)]
    _log("building the %s graph on %s" % (workload, dev))
    graph = build_graph(workload, dev)
    ds = D.get_dataset({"name": "SyntheticDataset", "device": dev, "graph": graph})
    _log("graph ready: %d users x %d items, %d train pairs" % (ds.n_users, ds.n_items, len(ds)))
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
            shard = DimShard(rank, world)
            shard_model_dims(m, shard)
            d = m.embedding_size
            if big and not args.no_hybrid:  # replicas of the column-shard group split the rows instead of repeating work
                partition = shard.row_partition(m.norm_adj, d)
    trainer_name = {"igcn": "IGCNTrainer", "sgl": "SGLTrainer", "half": "HALFTrainer"}.get(args.model, "BPRTrainer")
    tr = T.get_trainer({"name": trainer_name, "contrastive_reg": 0.1, "optimizer": "Adam", "lr": LR,
                        "l2_reg": 0.0 if args.model == "igcn" else L2_REG, "aux_reg": 0.01, "device": dev,
                        "n_epochs": 1, "batch_size": BATCH, "dataloader_num_workers": 0, "test_batch_size": 512,
                        "topks": TOPKS, "partition": partition}, ds, m)
    m.train()
    _log("model + trainer ready")
    eng = tr._engine()
    steps_per_epoch = tr.steps_per_epoch()
    K, W = args.steps, max(args.warmup, 3)
    l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
    # small shapes are L2-resident: flush between timed steps.  c4's tables (1.5 GB each) are 12x the L2: no flush needed.
    flush_buf = None if big else torch.empty(max(2 * l2_bytes, 256 << 20), dtype=torch.uint8, device=dev)

    def flush():
        if flush_buf is not None:
            flush_buf.fill_(1)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_steps(fn, n, do_flush):
        """device time of n calls of fn, CUDA events on the launching stream; L2 flushed (untimed) between calls"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            if do_flush:
                flush()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in ev]

    peak, peak_src = peaks()
    hbm_peak = float(peak["hbm_gbs"])
    n = ds.n_users + ds.n_items
    line = {}
    with ClockSampler(local) as clocks:
        if workload != "c5":
            for _ in range(W):
                eng.step()
            barrier()
            _log("warm-up done (%d steps), timing %d steps" % (W, K))
            launches0 = _abi.launch_count()
            times = timed_steps(eng.step, K, True)
            barrier()
            # the same K steps back to back (L2 warm: the steady state of a real epoch), one event pair
            t_warm = timed_steps(lambda: [eng.step() for _ in range(K)], 1, False)[0] / K
            total_ms = float(sum(times))
            if world > 1:
                t = torch.tensor([total_ms, t_warm], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                total_ms, t_warm = float(t[0]), float(t[1])
            ms_per_step = total_ms / K
            value = 1e3 / (ms_per_step * steps_per_epoch)
            kernels_per_step = eng.kernels_per_step()
            assert _abi.launch_count() >= launches0  # replays do not pass through the host counter

            _log("device-timed: %.3f ms/step" % ms_per_step)
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
                flush()
                eng.step(host_batch=b)
                loss = eng.last_loss()  # device->host read of the step's loss (the reference's loss.item(), trainer.py:428)
            t1.record()
            torch.cuda.synchronize()
            e2e_ms = max(t0.elapsed_time(t1), 1e3 * (time.perf_counter() - wall0)) / K
            # a flush inside this region cannot be hoisted out of a host-synchronous loop: subtract its measured cost
            flush_ms = float(np.median(timed_steps(flush, 10, False))) if flush_buf is not None else 0.0
            e2e_ms_net = max(e2e_ms - flush_ms, 1e-6)
            if world > 1:
                t = torch.tensor([e2e_ms_net], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                e2e_ms_net = float(t[0])
            e2e_value = 1e3 / (e2e_ms_net * steps_per_epoch)

            _log("e2e: %.3f ms/step" % e2e_ms_net)
            # ---- roofline of the dominant kernel: one SpMM layer as the step runs it (Y = A X, acc += Y) ----
            adj = m.norm_adj if partition is None else partition.local_op
            nnz_local = int(adj.item_end[:adj.n_items].sum() - adj.item_start[:adj.n_items].sum()) if partition is not None else adj.nnz
            rows_local = n if partition is None else partition.n_local_rows
            x = torch.randn((n, d), device=dev)
            y = torch.empty_like(x)
            acc = torch.zeros_like(x)
            spmm_call = lambda: ops.spmm(adj, x, y=y, addend=acc, out=acc)  # noqa: E731
            l0 = _abi.launch_count()
            spmm_call()
            launches_per_layer = _abi.launch_count() - l0
            for _ in range(2):
                spmm_call()
            reps = 10 if big else 30
            spmm_ms = float(np.median(timed_steps(spmm_call, reps, True)))
            spmm_ms_warm = timed_steps(lambda: [spmm_call() for _ in range(reps)], 1, False)[0] / reps
            del x, y, acc
            # algorithmic bytes per layer: nnz*(4+4) + (N+1)*4 + 2*N*D*4
            bytes_min = nnz_local * 8 + (rows_local + 1) * 4 + (n + rows_local) * d * 4
            bytes_gather = nnz_local * 8 + nnz_local * d * 4 + rows_local * d * 4
            achieved = bytes_min / (spmm_ms * 1e-3) / 1e9
            blk = adj.blocked_for(d) if partition is None else None
            roofline = {"bound": "hbm", "kernel": "spmm_items_kernel (one propagation layer, D=%d%s)" % (
                            d, "" if blk is None else ", column-blocked: %d passes x %d-column sweeps" % (blk[0].n_passes, blk[1])),
                        "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "traffic": None, "traffic_note": "not measurable inside the run; ncu dram__bytes per layer are in profiles/",
                        "peak_source": peak_src, "bytes_min": bytes_min, "launch_ms": spmm_ms,
                        "launches_per_layer": launches_per_layer, "launch_ms_l2_warm": spmm_ms_warm,
                        "effective_gather_gbs": bytes_gather / (spmm_ms * 1e-3) / 1e9,
                        "effective_gather_note": "nnz*D*4 gathered bytes / t; a pure random gather of whole rows (no index stream, "
                                                 "no FMA) reaches 17.5 TB/s from a table that fits L2 and 7.5 TB/s from a 1.5 GB one "
                                                 "(profiles/r02_l2_gather_ceiling.txt, DESIGN 4.1)"}
            line.update({"metric": "bpr_epochs_per_sec", "value": value, "unit": "epochs/s", "ms_per_step": ms_per_step,
                         "ms_per_step_l2_warm": t_warm, "epochs_per_sec_l2_warm": 1e3 / (t_warm * steps_per_epoch),
                         "e2e": {"value": e2e_value, "unit": "epochs/s", "h2d_bytes_per_step": BATCH * 3 * 8,
                                 "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms_net, "last_loss": loss},
                         "gpu_launches": kernels_per_step * K, "kernels_per_step": kernels_per_step, "roofline": roofline})

        # ---- full-rank evaluation (users/s): representation + fused score/mask/top-20 + metrics ----
        if workload != "c5":
            _log("SpMM layer: %.3f ms" % spmm_ms)
        eval_info = None
        if not args.no_eval:  # every rank takes part: the sweep is user-sharded and gathers the columns collectively
            m.eval()
            n_eval = 1 if big else 2
            tr.eval("test")
            barrier()
            _log("first evaluation pass done")
            w0 = time.perf_counter()
            for _ in range(n_eval):
                _, metrics, _ = tr.eval("test")
            barrier()
            eval_s = (time.perf_counter() - w0) / n_eval
            ke = timed_steps(lambda: tr.recommend_all("test"), 2 if big else 3, not big)
            ke_ms = float(min(ke))
            if world > 1:
                t = torch.tensor([eval_s, ke_ms], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                eval_s, ke_ms = float(t[0]), float(t[1])
            d_full = getattr(m, "full_embedding_size", m.embedding_size)
            score_flops = 2.0 * ds.n_users * ds.n_items * d_full
            tf_peak = float(peak.get("bf16_tflops_sustained" if big else "bf16_tflops", 1590.0))
            tf = score_flops / (ke_ms * 1e-3) / 1e12 / world  # per GPU: the sweep is user-sharded
            eval_info = {"users_per_s_e2e": ds.n_users / eval_s, "users_per_s_kernels": ds.n_users / (ke_ms * 1e-3),
                         "sweep_ms": ke_ms, "k": max(TOPKS), "recall@20": float(metrics["Recall"][20]),
                         "ndcg@20": float(metrics["NDCG"][20]), "precision": tr.eval_precision,
                         "trained_steps": eng.steps_done,
                         "score_roofline": {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s",
                                            "frac": tf / tf_peak, "note": "whole recommend_all sweep (propagation + prep + "
                                            "tcgen05 score/candidates + exact re-score) per GPU; see DESIGN 4.6"},
                         "includes": "trainer.eval('test'): get_rep + score/mask/top-K + fused rank-metrics pass + D2H of the metric sums"}
            if workload == "c5":
                line.update({"metric": "eval_users_per_sec", "value": ds.n_users / (ke_ms * 1e-3), "unit": "users/s",
                             "ms_per_step": ke_ms, "e2e": {"value": ds.n_users / eval_s, "unit": "users/s",
                                                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": (3 * len(TOPKS) + 1) * 8,
                                                           "note": "trainer.eval('test') end to end; inputs are the resident model"},
                             "gpu_launches": None, "roofline": eval_info["score_roofline"]})

    # ---- CPU baseline beside it (rank 0, N=1): bounded sample of the same workload on the host cores ----
    small = None
    if rank == 0 and world == 1 and big and not args.no_small:
        small = small_shape_leg(dev)
        _log("small-shape leg (c2): %.3f ms/step, evaluation %.2f ms" % (small["ms_per_step"], small["eval_ms"]))
    _log("GPU legs done")
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "lightgcn":
        port = CpuPort(graph, workload)
        _log("CPU port built (%d threads)" % port.threads)
        n_cpu = 1 if big else 3
        if not big:
            port.step()
        s_cpu = float(np.mean([port.step() for _ in range(n_cpu)]))
        cpu_baseline = {"value": 1.0 / (steps_per_epoch * s_cpu), "unit": "epochs/s", "cores": port.threads, "kind": "port",
                        "sample": port.sample_doc(n_cpu, steps_per_epoch), "ms_per_step": 1e3 * s_cpu}
    if rank == 0:
        out = {"metric": None, "value": None, "unit": None, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": None,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload + ": " + WORKLOAD_DOC[workload].replace("LightGCN", model_cfg["name"]), "batch": BATCH,
                          "steps_per_epoch": steps_per_epoch, "optimizer": "Adam", "sampler": "device (Philox)",
                          "graph_fingerprint": "%016x" % _fingerprint(graph),
                          "l2": ("inputs larger than L2 (tables of %d MB against %d MB of L2): no flush" % ((n * d * 4) >> 20, l2_bytes >> 20))
                          if big else "flushed between timed steps (write of %d MiB)" % (flush_buf.numel() >> 20),
                          "parallelism": "single GPU" if world == 1 else (
                              ("row-partitioned graph x%d, exchange fused into the SpMM / Adam epilogues (NVLink peer stores) "
                               "+ signal/wait hand-shake" % world) if args.parallelism == "peer" else
                              "row-partitioned graph x%d, per-layer NCCL block exchange" % world if args.parallelism == "row" else
                              ("embedding dimension sharded x%d (%d columns per GPU) x rows partitioned x%d among the GPUs that hold "
                               "the same columns (finished rows pushed over NVLink inside the SpMM / Adam epilogues), one [B,3] "
                               "all-reduce per step; evaluation user-sharded x%d" % (m._dim_shard.world, d, m._dim_shard.n_replicas, world))
                              if partition is not None else
                              "embedding dimension sharded x%d (%d columns per GPU) x %d replicas, one [B,3] all-reduce per "
                              "step; evaluation user-sharded x%d" % (m._dim_shard.world, d, world // m._dim_shard.world, world))}}
        out.update(line)
        out.update({"cpu_baseline": cpu_baseline, "eval": eval_info, "small_shape": small, "clocks": clocks.summary()})
        _emit(out)
    if world > 1:
        eng.close()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


_JSON_FD = None


def _claim_stdout():
    """stdout carries ONE JSON line: everything else written to descriptor 1 -- NCCL prints its version banner there from
    native code -- is sent to stderr, and the line goes out through a private duplicate of the original descriptor"""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the C2 leg reported beside the headline at N=1")
    ap.add_argument("--no-hybrid", action="store_true", help="multi-GPU: column shards only (narrow shards / identical replicas)")
    ap.add_argument("--model", default="lightgcn", choices=["lightgcn", "igcn", "mf", "sgl", "half"],
                    help="lightgcn is the headline; igcn = inductive template-feature layer + propagation (config.py:18-23)")
    ap.add_argument("--parallelism", default="dim", choices=["dim", "row", "peer"],
                    help="multi-GPU decomposition: embedding-dimension sharding (default), the row partition with NCCL "
                         "block exchange, or the row partition with the exchange fused into the kernels over peer memory")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU port)")
    run_own(args, rank, world)


if __name__ == "__main__":
    main()
