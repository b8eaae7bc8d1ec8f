#!/usr/bin/env python
"""Database-sharded search (BASELINE config 4) over N GPUs: parity with a single unsharded store + timing.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sharded_check.py [--records 1000000]

Every rank holds the replicated routing index and the store shard of a contiguous global-id range; DeviceShardedSearcher routes
query-parallel, all-gathers the candidate lists, refines its shard, all-gathers + merges the per-shard top-k (NCCL).  Rank 0 also
holds the whole store in a second context and checks that the sharded result is bit-identical to the unsharded search."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fspann_query_system_b200 import distributed as DD, hostsetup as HS, workloads as WL  # noqa: E402
from fspann_query_system_b200.gpu import GpuContext  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--records", type=int, default=0, help="override N")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--python-collectives", action="store_true", help="use distributed.DeviceShardedSearcher (torch.distributed all-gathers) "
                    "instead of fspann_sharded_search_batch_dev (NCCL inside the C library)")
    args = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cfg = WL.CONFIGS[args.config]
    if args.records:
        cfg = cfg.scaled(N=args.records)
    base = WL.base_vectors(cfg)
    alpha, r, omega = HS.build_gfunctions(base[:HS.MIN_SAMPLE_SIZE], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    gpu = GpuContext(lr)
    gpu.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
    gpu.routing_build(base, HS.staged_order(cfg.N), want_arrays=False)               # replicated routing index
    km = HS.KeyManager(WL.MASTER_KEY)
    gpu.keys_set(1, km.derive(1))
    lo, hi = DD.shard_range(cfg.N, rank, world)
    iv = WL.record_ivs(cfg.N, cfg.base_seed + 5)
    ct = gpu.encrypt_batch(np.arange(lo, hi, dtype=np.int32), base[lo:hi], iv[lo:hi], 1)
    gpu.store_upload(cfg.dim, iv[lo:hi], ct, np.ones(hi - lo, dtype=np.int32), id_base=lo, n_global=cfg.N)   # this rank's shard only
    if args.python_collectives:
        searcher = DD.DeviceShardedSearcher(gpu)
    else:                                                                            # the product path: ONE C-ABI call per batch
        ids = [gpu.comm_unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(ids, src=0)
        gpu.comm_init(world, rank, ids[0])

        class AbiSearcher:
            def __init__(self):
                self.stream = torch.cuda.ExternalStream(gpu.stream(), device=torch.device("cuda", lr))

            def search_batch_dev(self, dq, k, probes, hard_cap, B):
                Q = dq.shape[0]
                o = dict(top_ids=torch.empty((Q, k), dtype=torch.int32, device="cuda"), top_dist=torch.empty((Q, k), dtype=torch.float64, device="cuda"),
                         n_ret=torch.empty((Q,), dtype=torch.int32, device="cuda"), counters=torch.empty((Q, 6), dtype=torch.int64, device="cuda"))
                torch.cuda.current_stream().synchronize()
                gpu.sharded_search_batch_dev(Q, dq.data_ptr(), k, probes, hard_cap, B, 1, o["top_ids"].data_ptr(), o["top_dist"].data_ptr(),
                                             o["n_ret"].data_ptr(), o["counters"].data_ptr())
                return o
        searcher = AbiSearcher()
    batches = []
    for b in range(3):
        qcfg = cfg.scaled(name=cfg.name)
        object.__setattr__(qcfg, "query_seed", cfg.query_seed + 7919 * b)
        batches.append(torch.from_numpy(WL.query_vectors(qcfg)).cuda())
    for b in batches:                                                                # warm-up
        out = searcher.search_batch_dev(b, cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
    torch.cuda.synchronize()
    gpu.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.steps):
        out = searcher.search_batch_dev(batches[s % 3], cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
    torch.cuda.synchronize()
    gpu.sync()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    stage = None if args.python_collectives else dict(gpu.sharded_stage_ms(), refine_split=gpu.stage_ms())
    same = None
    if rank == 0:
        full = GpuContext(lr)
        full.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
        full.routing_build(base, HS.staged_order(cfg.N), want_arrays=False)
        full.keys_set(1, km.derive(1))
        full.store_upload(cfg.dim, iv, full.encrypt_batch(np.arange(cfg.N, dtype=np.int32), base, iv, 1), np.ones(cfg.N, dtype=np.int32))
        b = batches[(args.steps - 1) % 3]
        ref = full.search_batch(b.cpu().numpy(), cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
        same = bool(np.array_equal(out["top_ids"].cpu().numpy(), ref["top_ids"])
                    and np.array_equal(out["top_dist"].cpu().numpy().view(np.uint64), ref["top_dist"].view(np.uint64))
                    and np.array_equal(out["n_ret"].cpu().numpy(), ref["n_ret"]))
        print(json.dumps({"mode": "database-sharded (config 4 shape)", "n_gpus": world, "N": cfg.N, "dim": cfg.dim, "Q": cfg.Q, "B": cfg.B,
                          "ms_per_batch": 1e3 * dt / args.steps, "queries_per_s": cfg.Q * args.steps / dt,
                          "sharded_equals_unsharded": same, "collectives": "torch.distributed" if args.python_collectives else "C ABI (NCCL in libfspann_gpu.so)",
                          "stage_ms_rank0": stage, "timing": "wall clock around the steps (barrier + synchronize on both sides)"}),
              file=real_stdout, flush=True)
        full.close()
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
