"""Multi-GPU parity check, launched by tests/test_gpu_dist.py under torchrun (one rank per GPU, NCCL):
every decomposition must reproduce the single-GPU / reference results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.join(os.path.dirname(HERE), "inductive-recommendation_b200"), os.path.dirname(HERE)]
from conftest import MODEL_CFG, golden_model, load_golden  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    import trainer as T
    from b200rec import ops
    from b200rec.dist import DimShard, RowPartition, shard_model_dims
    topks = [1, 5, 10, 15, 20]
    engines = []

    def log(*a):
        print('[rank %d]' % rank, *a, flush=True)

    def trainer_for(g, name, ds, m, **extra):
        is_igcn = MODEL_CFG[name]["name"] in ("IGCN", "IMF")
        cfg = {"name": "IGCNTrainer" if is_igcn else "BPRTrainer", "optimizer": "Adam", "lr": float(g["lr"]),
               "l2_reg": float(g["l2_reg"]), "aux_reg": float(g["aux_reg"]), "device": dev, "n_epochs": 1,
               "batch_size": int(g["batch"].shape[0]), "dataloader_num_workers": 0, "test_batch_size": 128, "topks": topks}
        cfg.update(extra)
        return T.get_trainer(cfg, ds, m)

    # ---- embedding-dimension sharding: one fused step on the reference's recorded batch ----
    for name in ("lightgcn_tiny", "igcn_tiny", "mf_tiny", "lightgcn_d128"):
        g = load_golden(name)
        ds, m = golden_model(g, name, device=dev)
        shard = DimShard(rank, world)
        shard_model_dims(m, shard)
        tr = trainer_for(g, name, ds, m)
        log('dim', name, 'built')
        # evaluation before the step: user-sharded top-K ids equal the reference's
        rec = tr.recommend_all("test").cpu().numpy()
        ref_ids, ref_val = g["topk_ids_test"], g["topk_val_test"]
        sep = np.ones_like(ref_ids, dtype=bool)
        dd = np.abs(np.diff(ref_val, axis=1)) > 1e-6
        sep[:, 1:] &= dd
        sep[:, :-1] &= dd
        assert np.array_equal(rec[sep], ref_ids[sep]), name
        log('dim', name, 'eval ok')
        m.train()
        eng = tr._engine()
        engines.append(eng)
        hb = torch.from_numpy(g["batch"]).pin_memory()
        ha = torch.from_numpy(g["aux_batch"]).pin_memory() if "aux_batch" in g else None
        hk = None
        if "drop_keep" in g:
            keep = np.unpackbits(g["drop_keep"])[: int(g["drop_nnz"])].astype(bool)
            hk = ops.pack_keep_bits(torch.from_numpy(keep).to(dev))
        eng.step(host_batch=hb, host_aux_batch=ha, host_keep_bits=hk)
        log('dim', name, 'step done')
        assert abs(eng.last_loss() - float(g["loss"])) < 2e-6, (name, eng.last_loss(), float(g["loss"]))
        assert abs(eng.meter_avg() - float(g["loss"])) < 2e-6
        if name.startswith("mf"):
            full = shard.gather_cols(m._joint).cpu().numpy()
            ref = np.concatenate([g["user_emb1"], g["item_emb1"]], 0)
        else:
            full = shard.gather_cols(m.embedding.weight.data).cpu().numpy()
            ref = g["emb1"]
        np.testing.assert_allclose(full, ref, rtol=1e-5, atol=5e-6, err_msg=name)
        if "w1" in g:
            np.testing.assert_allclose(shard.gather_cols(m.w.data[None, :])[0].cpu().numpy(), g["w1"], rtol=1e-5, atol=5e-6)

    # ---- row partition: bit-identical to the single-rank propagation; engine step equals the reference ----
    for name in ("lightgcn_tiny", "lightgcn_d128"):
        g = load_golden(name)
        ds, m = golden_model(g, name, device=dev)
        L = int(g["n_layers"])
        x0 = m.embedding.weight.data
        single = torch.empty_like(x0)
        bufs = [torch.empty_like(x0), torch.empty_like(x0)]
        ops.propagate_fwd(m.norm_adj, x0, L, bufs, single)
        part = RowPartition(m.norm_adj, rank, world)
        multi = torch.empty_like(x0)
        part.propagate_fwd(m.norm_adj, x0, L, bufs, multi)
        assert torch.equal(single, multi), name + ": row-partitioned forward differs"
        log('row', name, 'fwd ok')
        gsrc = torch.randn(x0.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        d1, d2 = torch.empty_like(x0), torch.empty_like(x0)
        ops.propagate_bwd(m.norm_adj, gsrc, L, bufs, d1)
        part.propagate_bwd(m.norm_adj, gsrc, L, bufs, d2)
        for lo, hi in part.my_blocks:
            assert torch.equal(d1[lo:hi], d2[lo:hi]), name + ": row-partitioned backward differs"
        tr = trainer_for(g, name, ds, m, partition=part)
        m.train()
        eng = tr._engine()
        engines.append(eng)
        log('row', name, 'bwd ok')
        eng.step(host_batch=torch.from_numpy(g["batch"]).pin_memory())
        log('row', name, 'step done')
        assert abs(eng.last_loss() - float(g["loss"])) < 2e-6
        np.testing.assert_allclose(m.embedding.weight.data.cpu().numpy(), g["emb1"], rtol=1e-5, atol=5e-6)
    # ---- peer-memory row partition: the exchange is fused into the SpMM / Adam epilogues (NVLink stores) ----
    if os.environ.get("B200REC_SKIP_PEER", "0") != "1":
        from b200rec.dist import PeerRowPartition
        for name in ("lightgcn_tiny", "lightgcn_d128"):
            g = load_golden(name)
            ds, m = golden_model(g, name, device=dev)
            L = int(g["n_layers"])
            x0 = m.embedding.weight.data.clone()
            single = torch.empty_like(x0)
            bufs = [torch.empty_like(x0), torch.empty_like(x0)]
            ops.propagate_fwd(m.norm_adj, x0, L, bufs, single)
            part = PeerRowPartition(m.norm_adj, rank, world, x0.shape[1])
            log('peer', name, 'tables mapped')
            part.table.copy_(x0)
            torch.cuda.synchronize(); dist.barrier()
            part.propagate_fwd(m.norm_adj, part.table, L, None, part.rep)
            torch.cuda.synchronize()
            assert torch.equal(single, part.rep), name + ": peer row-partitioned forward differs"
            log('peer', name, 'fwd ok')
            gsrc = torch.randn(x0.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
            d1, d2 = torch.empty_like(x0), torch.empty_like(x0)
            ops.propagate_bwd(m.norm_adj, gsrc, L, bufs, d1)
            part.propagate_bwd(m.norm_adj, gsrc, L, None, d2)
            torch.cuda.synchronize()
            for lo, hi in part.my_blocks:
                assert torch.equal(d1[lo:hi], d2[lo:hi]), name + ": peer row-partitioned backward differs"
            log('peer', name, 'bwd ok')
            tr = trainer_for(g, name, ds, m, partition=part)
            m.train()
            eng = tr._engine()
            engines.append(eng)
            for use_graph in (False, True):
                with torch.no_grad():
                    m.embedding.weight.data.copy_(torch.from_numpy(g["emb0"]).to(dev))
                    eng.m.zero_(); eng.v.zero_(); eng.adam_step.zero_()
                torch.cuda.synchronize(); dist.barrier()
                eng.use_graph = use_graph
                eng.step(host_batch=torch.from_numpy(g["batch"]).pin_memory())
                torch.cuda.synchronize()
                assert abs(eng.last_loss() - float(g["loss"])) < 2e-6
                np.testing.assert_allclose(m.embedding.weight.data.cpu().numpy(), g["emb1"], rtol=1e-5, atol=5e-6)
                log('peer', name, 'engine step ok, graph =', use_graph)
    # ---- hybrid: column shards x row-partitioned replicas (bench.py's layout on 8 GPUs: 4 x 32 columns x 2 row halves) ----
    if world >= 4 and world % 2 == 0 and os.environ.get("B200REC_SKIP_PEER", "0") != "1":
        for name in ("lightgcn_d128", "lightgcn_tiny"):
            g = load_golden(name)
            ds, m = golden_model(g, name, device=dev)
            d_full = m.embedding_size
            shard = DimShard(rank, world, min_cols=d_full // (world // 2))
            shard_model_dims(m, shard)
            assert shard.world == world // 2 and shard.n_replicas == 2 and m.embedding_size == d_full // shard.world
            part = shard.row_partition(m.norm_adj, m.embedding_size)
            assert part is not None and part.world == 2 and part.rank == rank // shard.world
            log('hybrid', name, 'tables mapped')
            tr = trainer_for(g, name, ds, m, partition=part)
            rec = tr.recommend_all("test").cpu().numpy()
            ref_ids, ref_val = g["topk_ids_test"], g["topk_val_test"]
            sep = np.ones_like(ref_ids, dtype=bool)
            dd = np.abs(np.diff(ref_val, axis=1)) > 1e-6
            sep[:, 1:] &= dd
            sep[:, :-1] &= dd
            assert np.array_equal(rec[sep], ref_ids[sep]), name
            m.train()
            eng = tr._engine()
            engines.append(eng)
            lo, hi = shard.cols(d_full)
            for use_graph in (False, True):
                with torch.no_grad():
                    m.embedding.weight.data.copy_(torch.from_numpy(g["emb0"][:, lo:hi]).to(dev))
                    eng.m.zero_(); eng.v.zero_(); eng.adam_step.zero_(); eng.loss_accum.zero_()
                torch.cuda.synchronize(); dist.barrier()
                eng.use_graph = use_graph
                eng.step(host_batch=torch.from_numpy(g["batch"]).pin_memory())
                torch.cuda.synchronize()
                assert abs(eng.last_loss() - float(g["loss"])) < 2e-6, (name, eng.last_loss(), float(g["loss"]))
                full = shard.gather_cols(m.embedding.weight.data).cpu().numpy()
                np.testing.assert_allclose(full, g["emb1"], rtol=1e-5, atol=5e-6, err_msg=name)
                log('hybrid', name, 'engine step ok, graph =', use_graph)
    dist.barrier()
    for e in engines:
        e.close()
    del engines
    torch.cuda.synchronize()
    if rank == 0:
        print("DIST_GPU_CHECK_OK world=%d" % world, flush=True)
    dist.destroy_process_group()
    log("exit")


if __name__ == "__main__":
    main()
