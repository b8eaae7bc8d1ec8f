#!/usr/bin/env python
"""BASELINE config 4 at its stated size: Deep100M-shape synthetic 100 M x 96, routing index replicated, encrypted store sharded by id
range over the GPUs of one box, one 10 k-query batch through fspann_sharded_search_batch_dev (NCCL all-gathers inside the C library).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/c4_scale.py --records 100000000

The base set never exists on the host: every rank generates it on its GPU in 1 M-row chunks (torch Philox generator seeded per chunk ->
N(0,1) float32, L2-normalised, widened to FP64 like the reference's loaders), codes each chunk into the routing build
(fspann_routing_build_begin / add_dev / finish) and encrypts the chunks of its own id range straight into its HBM shard
(fspann_store_alloc_shard / fspann_store_encrypt_dev).  The reference itself stops at 75 M records (JVM heap, README.md:316-320).

Checks that do not need a reference run at this size (the oracle is not used here):
  * every rank returns the same result (checksums all-gathered);
  * exactness of the refined distances: for a sample of queries the returned vectors are regenerated and the sequential FP64 L2 is
    recomputed on the host -- it must equal the returned distance bit for bit (decryption + distance + merge at scale);
  * recall@10 against exact ground truth for the sample (brute force over the regenerated chunks, FP64).
With --verify (N small enough for the host) the same data is also pushed through the ordinary host-array path on rank 0 and the
sharded result must equal the unsharded fspann_search_batch result bit for bit.
"""
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


def gen_chunk(c: int, rows: int, dim: int, seed: int) -> torch.Tensor:
    """Rows [c*chunk, c*chunk + rows) of the base set, FP64 on the current CUDA device (float32 values widened)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed + c)
    x = torch.randn((rows, dim), generator=g, device="cuda", dtype=torch.float32)
    x = x / torch.linalg.vector_norm(x, dim=1, keepdim=True)
    return x.to(torch.float64).contiguous()


def gen_ivs(c: int, rows: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cuda")
    g.manual_seed(seed * 7919 + c)
    return torch.randint(0, 256, (rows, 12), generator=g, device="cuda", dtype=torch.uint8).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--chunk", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--sample", type=int, default=64, help="queries whose results are re-derived exactly")
    ap.add_argument("--verify", action="store_true", help="also run the host-array path on rank 0 and compare (small N only)")
    args = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cfg = WL.C4.scaled(N=args.records)
    N, dim, k, chunk = cfg.N, cfg.dim, cfg.k, args.chunk
    n_chunks = (N + chunk - 1) // chunk
    lo, hi = DD.shard_range(N, rank, world)
    t_all = time.time()

    def log(*a):
        if rank == 0:
            print("[c4]", *a, file=sys.stderr, flush=True)

    first = gen_chunk(0, min(chunk, N), dim, cfg.base_seed)
    alpha, r, omega = HS.build_gfunctions(first[:HS.MIN_SAMPLE_SIZE].cpu().numpy(), cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    gpu = GpuContext(lr)
    gpu.gfunctions_upload(dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
    km = HS.KeyManager(WL.MASTER_KEY)
    gpu.keys_set(1, km.derive(1))
    gpu.store_alloc_shard(dim, lo, hi - lo, N)
    gpu.routing_build_begin(N)
    queries = WL.query_vectors(cfg)
    Q = queries.shape[0]
    ns = min(args.sample, Q)
    dq_s = torch.from_numpy(queries[:ns]).cuda()
    gt_d = torch.full((ns, k), float("inf"), dtype=torch.float64, device="cuda")
    gt_i = torch.full((ns, k), -1, dtype=torch.int64, device="cuda")
    t0 = time.time()
    host_chunks = []
    for c in range(n_chunks):
        rows = min(chunk, N - c * chunk)
        x = first if c == 0 else gen_chunk(c, rows, dim, cfg.base_seed)
        torch.cuda.synchronize()
        gpu.routing_build_add_dev(c * chunk, rows, x.data_ptr())                      # every rank codes the whole base set (replicated index)
        a, b = max(lo, c * chunk), min(hi, c * chunk + rows)
        if b > a:                                                                     # rows of my id range: encryptToPoint into my shard
            ivs = gen_ivs(c, rows, cfg.base_seed)
            torch.cuda.synchronize()
            gpu.store_encrypt_dev(a, b - a, x.data_ptr() + (a - c * chunk) * dim * 8, ivs.data_ptr() + (a - c * chunk) * 12, 1)
        if rank == 0:                                                                 # exact ground truth for the sample, chunk by chunk
            d2 = (dq_s * dq_s).sum(1, keepdim=True) - 2.0 * dq_s @ x.T + (x * x).sum(1)[None, :]
            cd, ci = torch.topk(d2, k, dim=1, largest=False)
            md, mi = torch.cat([gt_d, cd], 1), torch.cat([gt_i, ci + c * chunk], 1)
            o = torch.argsort(md, dim=1, stable=True)[:, :k]
            gt_d, gt_i = torch.gather(md, 1, o), torch.gather(mi, 1, o)
        if args.verify and rank == 0:
            host_chunks.append((x.cpu().numpy(), gen_ivs(c, rows, cfg.base_seed).cpu().numpy()))
        gpu.sync()
        del x
    t_code = time.time() - t0
    t0 = time.time()
    gpu.routing_build_finish(None)
    t_part = time.time() - t0
    treeified = gpu.get_info("build_treeified")
    log(f"N={N}: coding + shard encryption {t_code:.1f}s, partition build {t_part:.1f}s, build_treeified={treeified}, "
        f"HBM in use {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB free of {torch.cuda.mem_get_info()[1] / 2**30:.1f}")

    cid = [gpu.comm_unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(cid, src=0)
    gpu.comm_init(world, rank, cid[0])
    batches = []
    for b in range(3):
        qcfg = cfg.scaled(name=cfg.name)
        object.__setattr__(qcfg, "query_seed", cfg.query_seed + 7919 * b)
        batches.append(torch.from_numpy(WL.query_vectors(qcfg)).cuda())
    ids = torch.empty((Q, k), dtype=torch.int32, device="cuda"); dd = torch.empty((Q, k), dtype=torch.float64, device="cuda")
    nr = torch.empty((Q,), dtype=torch.int32, device="cuda"); cn = torch.empty((Q, 6), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def step(i):
        gpu.sharded_search_batch_dev(Q, batches[i % 3].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 1, ids.data_ptr(), dd.data_ptr(), nr.data_ptr(), cn.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); gpu.sync()
    for i in range(3):
        step(i)
    barrier()
    stream = torch.cuda.ExternalStream(gpu.stream(), device=torch.device("cuda", lr))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    step(0)
    barrier()
    stage = dict(gpu.sharded_stage_ms(), refine_split={kk: vv for kk, vv in gpu.stage_ms().items() if kk in ("group", "verify", "decrypt", "topk")})
    h_ids, h_dd, h_nr, h_cn = ids.cpu().numpy(), dd.cpu().numpy(), nr.cpu().numpy(), cn.cpu().numpy()
    # (1) every rank holds the same result
    chk = torch.tensor([int(h_ids.astype(np.int64).sum()), int(h_dd.view(np.int64).sum() & 0x7FFFFFFFFFFF), int(h_nr.sum())], dtype=torch.int64, device="cuda")
    if world > 1:
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same_everywhere = all(bool((c == chk).all().item()) for c in allc)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    else:
        same_everywhere = True
    out = None
    if rank == 0:
        # (2) the returned distances, re-derived: regenerate each returned vector and recompute the sequential FP64 L2 (QSI:364-372)
        want = sorted({int(i) for i in h_ids[:ns].ravel() if i >= 0})
        exact, total = 0, 0
        rows = {}                                                                      # only the returned rows leave the device
        for c in sorted({i // chunk for i in want}):
            x = gen_chunk(c, min(chunk, N - c * chunk), dim, cfg.base_seed)
            sel = [i for i in want if i // chunk == c]
            got = x[torch.tensor([i % chunk for i in sel], device="cuda")].cpu().numpy()
            rows.update({i: got[n_] for n_, i in enumerate(sel)})
            del x
        q0 = batches[0][:ns].cpu().numpy()
        for q in range(ns):
            for j in range(int(h_nr[q])):
                i = int(h_ids[q, j]); v = rows[i]
                s = 0.0
                for t_ in range(dim):
                    d = q0[q, t_] - v[t_]
                    s = s + d * d
                total += 1
                exact += int(np.float64(np.sqrt(s)).view(np.uint64) == h_dd[q, j].view(np.uint64))
        # (3) recall@10 of the sample against exact ground truth
        gt = gt_i.cpu().numpy()
        rec = float(np.mean([len(set(gt[q].tolist()) & set(h_ids[q, :int(h_nr[q])].tolist())) / k for q in range(ns)]))
        verify = None
        if args.verify:
            base = np.concatenate([h[0] for h in host_chunks]); iv = np.concatenate([h[1] for h in host_chunks])
            full = GpuContext(lr)
            full.gfunctions_upload(dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
            full.routing_build(base, HS.staged_order(N), want_arrays=False)
            full.keys_set(1, km.derive(1))
            full.store_upload(dim, iv, full.encrypt_batch(np.arange(N, dtype=np.int32), base, iv, 1), np.ones(N, dtype=np.int32))
            ref = full.search_batch(batches[0].cpu().numpy(), k, cfg.probes, cfg.hard_cap, cfg.B)
            verify = bool(np.array_equal(ref["top_ids"], h_ids) and np.array_equal(ref["top_dist"].view(np.uint64), h_dd.view(np.uint64))
                          and np.array_equal(ref["n_ret"], h_nr) and np.array_equal(ref["counters"], h_cn))
            full.close()
        pairs = int(h_cn[:, 5].sum())
        out = {"mode": "BASELINE config 4 (database-sharded, NCCL inside the C library)", "n_gpus": world, "N": N, "dim": dim, "Q": Q, "B": cfg.B, "k": k,
               "records_per_shard": hi - lo, "ms_per_batch": ms, "queries_per_s": Q / (ms * 1e-3), "stage_ms_rank0": stage,
               "same_result_on_every_rank": same_everywhere, "returned_distances_exact": f"{exact}/{total}", "recall_at_10_sample": rec, "sample_queries": ns,
               "mean_returned": float(h_nr.mean()), "pairs": pairs, "equals_unsharded_host_path": verify, "build_treeified": int(treeified),
               "setup_s": {"coding_and_shard_encryption": t_code, "partition_build": t_part, "total_wall": time.time() - t_all},
               "reference_ceiling": "75 M records (JVM heap exhaustion at 92.6 M, README.md:316-320)",
               "timing": "CUDA events on the library stream around the steps, max over ranks"}
        print(json.dumps(out), file=real_stdout, flush=True)
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
