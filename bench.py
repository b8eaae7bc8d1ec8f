#!/usr/bin/env python
"""bench.py -- queries/sec of the FSPANN query hot path (TokenGen -> Route -> Refine) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

A "step" = one pass of the hot path over one batch of synthetic queries (BASELINE.json configs[1]: SIFT1M-shape
1M x 128, L=8 tables, D=8, m=24, lambda=2, B=1024, k=10, 10k-query batch).  For N>1 every rank holds a replica of the
routing index and the record store and searches its own 10k batch per step (weak scaling, no data-path collective).
`value`  = queries/sec with the query batches already resident in HBM (CUDA events on the library's stream).
`e2e`    = the same through the host-buffer C-ABI call (pinned host queries -> H2D -> search -> D2H of results).
Further legs of the same JSON line (all outside the timed regions of value / e2e):
  roofline.stage_frac / reuse_factor / no_reuse : the whole refine stage (group + verify + decrypt + top-k) against the HBM roofline, how
                      many (query, candidate) pairs share one decrypted record, and the same stage with NO reuse (every pair names a
                      different record);
  sharded           : ONE batch over the store sharded by id range across the N GPUs through fspann_sharded_search_batch_dev (NCCL
                      all-gathers inside the C library) -- strong scaling of config 4's split, checked against the unsharded result;
  strong_replicated : ONE batch split N ways over replicated state (config 3's split) -- strong scaling without any collective;
  configs           : (N=1) BASELINE configs 1, 3, the config-4 shape at a single-GPU size and the reference's published operating point
                      SIFT_P6_BALANCED, each with a sample checked against the oracle.
The oracle (oracle/) is used ONLY for the cpu_baseline / --impl reference legs (as the thing being timed) and as the checker of samples.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fspann_query_system_b200 import hostsetup as HS  # noqa: E402
from fspann_query_system_b200 import workloads as WL  # noqa: E402


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


def load_traffic(workload_is_c2: bool):
    """DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture (C2 only), else None."""
    if not workload_is_c2:
        return None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_decrypt_traffic.json")))
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions: NVML from a thread every 20 ms, started before the warm-up steps (more often perturbs the host-buffer leg: NVML queries contend with the driver; the timed regions last tens of
    milliseconds, too short for `nvidia-smi -lms`); falls back to one nvidia-smi query if NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu, self.sm, self.mask, self.max_mhz, self.stop_flag, self.t, self.h = gpu_index, [], 0, None, False, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: match by PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(hh).bus) == int(bus):
                        self.h = hh
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv, self.h = None, None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.020)

    def start(self):
        if self.h is not None:
            self.stop_flag = False
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def pause(self):
        if self.t is not None:
            self.stop_flag = True
            self.t.join()
            self.t = None

    def stop(self):
        self.pause()
        if self.h is None:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": ["sampled once after the run (NVML unavailable)"], "samples": 1}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"], "samples": 0}
        reasons = [n for bit, n in self.REASONS.items() if self.mask & bit]
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm)}


def build_world(cfg: WL.Config, gpu, n_batches: int, rank: int, quiet: bool = False):
    """Setup (host side, untimed): synthetic base set, GFunctions, encrypted store, greedy partitions; upload to HBM."""
    t0 = time.time()
    log_ = (lambda *a: None) if quiet else log
    base = WL.base_vectors(cfg)
    alpha, r, omega = HS.build_gfunctions(base[:HS.MIN_SAMPLE_SIZE], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    log_(f"base {base.shape} + GFunctions in {time.time() - t0:.1f}s")
    # Setup on the device: coding of the base set + GreedyPartitioner.build per division, then bulk encryptToPoint
    setup = {}
    t0 = time.time()
    gpu.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
    mn, mx, rep, ids = gpu.routing_build(base, HS.staged_order(cfg.N))
    setup["coding_and_partition_build_s"] = time.time() - t0
    log_(f"base codes + greedy partitions ({cfg.T * cfg.D} x {mn.shape[1]}) on the device in {time.time() - t0:.1f}s (incl. H2D of the base set and D2H of the index)")
    t0 = time.time()
    km = HS.KeyManager(WL.MASTER_KEY)
    iv = WL.record_ivs(cfg.N, cfg.base_seed + 5)
    gpu.keys_set(1, km.derive(1))
    ct = gpu.encrypt_batch(np.arange(cfg.N, dtype=np.int32), base, iv, 1)
    ver = np.ones(cfg.N, dtype=np.int32)
    setup["encrypt_s"] = time.time() - t0
    gpu.store_upload(cfg.dim, iv, ct, ver)
    log_(f"encrypted store ({ct.nbytes / 1e9:.2f} GB) on the device in {setup['encrypt_s']:.1f}s")
    batches = make_batches(cfg, n_batches, rank)
    world = dict(alpha=alpha, r=r, omega=omega, mn=mn, mx=mx, rep=rep, ids=ids, iv=iv, ct=ct, ver=ver, keys={1: km.derive(1)}, base=base, setup=setup)
    return world, batches


def make_batches(cfg, n_batches: int, rank: int):
    out = []
    for b in range(n_batches):
        qcfg = cfg.scaled(name=cfg.name)
        object.__setattr__(qcfg, "query_seed", cfg.query_seed + 7919 * b + 104729 * rank)
        out.append(WL.query_vectors(qcfg))
    return out


class DevRunner:
    """Device-resident buffers + timing of K steps of one search entry point on the library's stream."""

    def __init__(self, gpu, torch, local_rank, Q, k):
        self.gpu, self.torch, self.Q, self.k = gpu, torch, Q, k
        self.stream = torch.cuda.ExternalStream(gpu.stream(), device=torch.device("cuda", local_rank))
        self.ids = torch.empty((Q, k), dtype=torch.int32, device="cuda")
        self.dist = torch.empty((Q, k), dtype=torch.float64, device="cuda")
        self.nret = torch.empty((Q,), dtype=torch.int32, device="cuda")
        self.cnt = torch.empty((Q, 6), dtype=torch.int64, device="cuda")

    def timed(self, fn, K, Wm, barrier):
        torch = self.torch
        for i in range(Wm):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for i in range(K):
            fn(Wm + i)
        e1.record(self.stream)
        barrier()
        return e0.elapsed_time(e1)

    def result(self):
        t = self.torch
        t.cuda.synchronize(); self.gpu.sync()
        return dict(top_ids=self.ids.cpu().numpy(), top_dist=self.dist.cpu().numpy(), n_ret=self.nret.cpu().numpy(), counters=self.cnt.cpu().numpy())


def same_result(a, b, counters=True):
    ok = np.array_equal(a["top_ids"], b["top_ids"]) and np.array_equal(a["top_dist"].view(np.uint64), b["top_dist"].view(np.uint64)) and \
        np.array_equal(a["n_ret"], b["n_ret"])
    return bool(ok and (not counters or np.array_equal(a["counters"], b["counters"])))


def oracle_sample_check(cfg, world, gpu, queries, nq, threads):
    """First nq queries through the oracle (all host threads) vs the GPU: ids exact, FP64 distances bit-exact, counters equal."""
    O, ix, st = oracle_objects(cfg, world)
    dt, ref = cpu_run(cfg, O, ix, st, queries[:nq], threads)
    got = gpu.search_batch(queries[:nq], cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
    return same_result(got, ref), nq / dt


def run_extra_config(name, cfg, torch, local_rank, steps, warmup, threads, log_fn):
    """One of BASELINE's other configs on one GPU: Setup on the device, `steps` timed batches, a sample checked against the oracle."""
    from fspann_query_system_b200.gpu import GpuContext
    t0 = time.time()
    gpu = GpuContext(local_rank)
    try:
        world, batches = build_world(cfg, gpu, min(steps + warmup, 3), 0, quiet=True)
        Q, k = cfg.Q, cfg.k
        run = DevRunner(gpu, torch, local_rank, Q, k)
        d_b = [torch.from_numpy(b).cuda() for b in batches]
        torch.cuda.synchronize()

        def barrier():
            torch.cuda.synchronize(); gpu.sync()

        def step(i):
            gpu.search_batch_dev(Q, d_b[i % len(d_b)].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 0, 1, run.ids.data_ptr(), run.dist.data_ptr(),
                                 run.nret.data_ptr(), run.cnt.data_ptr())
        ms = run.timed(step, steps, warmup, barrier)
        step(0)
        stage = {kk: vv for kk, vv in gpu.stage_ms().items() if kk != "launches"}
        res = run.result()
        pairs = int(res["counters"][:, 5].sum())
        touched = int(gpu.touched(clear=True).size)
        stage_ms = stage["group"] + stage["verify"] + stage["decrypt"] + stage["topk"]
        alg = pairs * (8 * cfg.dim + 36) + Q * (8 * cfg.dim + 12 * k)
        peak, _ = load_peaks()
        nq = min(Q, 64 if cfg.B > 4096 or cfg.N > 2_000_000 else 256)
        ok, cpu_qps = oracle_sample_check(cfg, world, gpu, batches[0], nq, threads)
        out = {"workload": f"{cfg.name}: N={cfg.N} d={cfg.dim} T={cfg.T} D={cfg.D} m={cfg.m} probes={cfg.probes} B={cfg.B} k={cfg.k} cap={cfg.hard_cap} Q={Q}",
               "queries_per_s": Q * steps / (ms * 1e-3), "ms_per_batch": ms / steps, "stage_ms": stage, "route_path": {1: "fast", 2: "general"}.get(gpu.get_info("last_route_path"), "?"),
               "route_queries_handed_to_fallback": gpu.get_info("route_overflowed"), "hard_cap_binds": bool(cfg.hard_cap <= cfg.T * cfg.D * cfg.probes * 64 - 64),
               "refine_stage_frac_of_hbm": alg / (stage_ms * 1e-3) / 1e9 / peak if stage_ms > 0 else None,
               "reuse_factor": pairs / max(touched, 1), "mean_returned": float(res["n_ret"].mean()), "retried_queries": int(res["counters"][:, 4].sum()),
               "gpu_matches_oracle_on_sample": ok, "oracle_sample_queries": nq, "cpu_port_queries_per_s": cpu_qps, "cpu_threads": threads,
               "setup_on_device_s": world["setup"], "leg_wall_s": None}
        out["leg_wall_s"] = time.time() - t0
        log_fn(f"config {name}: {out['queries_per_s']:.0f} q/s, {out['ms_per_batch']:.2f} ms/batch, route={out['route_path']}, oracle sample ok={ok}, {out['leg_wall_s']:.1f}s")
        return out
    finally:
        gpu.close()


def oracle_objects(cfg, world):
    from oracle import oracle as O
    g = O.GFunctions(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, world["alpha"], world["r"], world["omega"])
    ix = O.Index(g, cfg.N, world["mn"].shape[1], world["mn"], world["mx"], world["rep"], world["ids"])
    st = O.Store(cfg.dim, world["iv"], world["ct"], world["ver"], dict(world["keys"]))
    return O, ix, st


def cpu_run(cfg, O, ix, st, queries, threads: int):
    """The reference algorithm (oracle port) over `queries`, statically partitioned over `threads` host threads."""
    Q = queries.shape[0]
    out = dict(top_ids=np.full((Q, cfg.k), -1, dtype=np.int32), top_dist=np.full((Q, cfg.k), np.nan), n_ret=np.zeros(Q, dtype=np.int32),
               counters=np.zeros((Q, 6), dtype=np.int64))
    bounds = np.linspace(0, Q, threads + 1).astype(int)
    ts = [threading.Thread(target=O.search_batch, args=(ix, st, queries, cfg.k, cfg.probes, cfg.hard_cap, cfg.B, 0, int(bounds[i]), int(bounds[i + 1]), out))
          for i in range(threads) if bounds[i + 1] > bounds[i]]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, out


def world_cpu_only(cfg):
    """--impl reference needs the same world without touching our CUDA library: the base codes come from the oracle."""
    from oracle import oracle as O
    base = WL.base_vectors(cfg)
    alpha, r, omega = HS.build_gfunctions(base[:HS.MIN_SAMPLE_SIZE], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    g = O.GFunctions(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
    threads = os.cpu_count() or 1
    codes = np.zeros((cfg.N, cfg.T * cfg.D, cfg.W), dtype=np.uint64)
    bounds = np.linspace(0, cfg.N, threads * 4 + 1).astype(int)

    def work(i):
        codes[bounds[i]:bounds[i + 1]] = O.tokengen_batch(base[bounds[i]:bounds[i + 1]], g)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads * 4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    mn, mx, rep, ids = HS.build_partitions(codes, HS.staged_order(cfg.N))
    km = HS.KeyManager(WL.MASTER_KEY)
    iv = WL.record_ivs(cfg.N, cfg.base_seed + 5)
    ct = O.encrypt_store(base, 1, km.derive(1), iv)
    return dict(alpha=alpha, r=r, omega=omega, mn=mn, mx=mx, rep=rep, ids=ids, iv=iv, ct=ct, ver=np.ones(cfg.N, dtype=np.int32), keys={1: km.derive(1)})


def main():
    sys.setswitchinterval(1e-4)     # the clock-sampling thread must hand the GIL back at once when a synchronous search call returns
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("FSPANN_BENCH_CONFIG", "C2"))
    ap.add_argument("--n", type=int, default=0, help="override N (debug)")
    ap.add_argument("--q", type=int, default=0, help="override batch size (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mixed", action="store_true", help="skip the config-5 Rotate + Migrate leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the legs for BASELINE configs 1 / 3 / 4-shape and the published operating point")
    ap.add_argument("--no-sharded", action="store_true", help="skip the database-sharded and strong-scaling legs")
    ap.add_argument("--extras", default="C1,C3,C4s,P6,SUB1", help="which extra configs to run at N=1")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = WL.CONFIGS[args.config]
    if args.n or args.q:
        cfg = cfg.scaled(N=args.n or None, Q=args.q or None)
    metric = "queries/sec at recall@10 (fixed budget B)"
    config = {"workload": f"{cfg.name}: N={cfg.N} d={cfg.dim} T={cfg.T} D={cfg.D} m={cfg.m} lambda={cfg.lam} probes={cfg.probes} "
                          f"B={cfg.B} k={cfg.k} maxGlobalCandidates={cfg.max_global} Q={cfg.Q}/step/GPU",
              "parallelism": f"query batch per GPU x{max(args.gpus, 1)}, routing index + store replicated, no collective",
              "l2": "inputs larger than L2 (store 1.06 GB + routing ids 0.26 GB vs 126 MB L2); a different query batch every step"}

    # ------------------------------------------------------------------ reference arm: oracle port on the host cores
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle as O
        threads = os.cpu_count() or 1
        world = world_cpu_only(cfg)
        O, ix, st = oracle_objects(cfg, world)
        per_step = max(threads * 48, 64)
        times = []
        for s in range(args.warmup + args.steps):
            qcfg = cfg.scaled(name=cfg.name)
            object.__setattr__(qcfg, "query_seed", cfg.query_seed + 7919 * s)
            qs = WL.query_vectors(qcfg, per_step)
            dt, _ = cpu_run(cfg, O, ix, st, qs, threads)
            if s >= args.warmup:
                times.append(dt)
        total = sum(times)
        val = per_step * args.steps / total
        sample = f"{per_step} queries/step of the {cfg.Q}-query workload, all {threads} host threads, in-memory store, OpenSSL AES-NI GCM"
        print(json.dumps({"impl": "reference", "metric": metric, "value": val, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "note": "reference is pure Java and no JVM exists in this image: this is the C restatement of its algorithm (oracle/), optimistic for the CPU"}))
        return

    # ------------------------------------------------------------------ our arm
    # stdout carries exactly ONE JSON line: anything a library prints there (e.g. NCCL's version banner) is sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the FSPANN hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from fspann_query_system_b200.gpu import GpuContext
    gpu = GpuContext(local_rank)
    K, Wm = args.steps, args.warmup
    n_batches = min(K + Wm, 6)
    world, batches = build_world(cfg, gpu, n_batches, rank)
    Q, k, dim = cfg.Q, cfg.k, cfg.dim

    stream = torch.cuda.ExternalStream(gpu.stream(), device=torch.device("cuda", local_rank))
    d_batches = [torch.from_numpy(b).cuda() for b in batches]
    d_ids = torch.empty((Q, k), dtype=torch.int32, device="cuda")
    d_dist = torch.empty((Q, k), dtype=torch.float64, device="cuda")
    d_nret = torch.empty((Q,), dtype=torch.int32, device="cuda")
    d_cnt = torch.empty((Q, 6), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def step_dev(i):
        gpu.search_batch_dev(Q, d_batches[i % n_batches].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 0, 1, d_ids.data_ptr(), d_dist.data_ptr(),
                             d_nret.data_ptr(), d_cnt.data_ptr())

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gpu.sync()

    # ---- value: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()                                                       # started before the warm-up: NVML's first queries are slow and stall submissions
    for i in range(Wm):
        step_dev(i)
    barrier()
    l0 = gpu.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step_dev(Wm + i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = gpu.launch_count() - l0
    sampler.pause()

    # ---- roofline of the dominant kernel (refine decrypt+distance), measured live with CUDA events on the launching stream
    stage = dict(tokengen=0.0, route=0.0, group=0.0, verify=0.0, decrypt=0.0, topk=0.0)
    pairs = 0
    for i in range(K):
        step_dev(Wm + i)
        sm = gpu.stage_ms()
        for key in stage:
            stage[key] += sm[key]
        pairs += int(d_cnt[:, 5].sum().item())
    dec_ms = stage["decrypt"] / K
    alg_bytes = pairs / K * (8 * dim + 36) + Q * (8 * dim + 12 * k)       # per pair 8d+36 B, per query 8d+12k B
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else 0.0
    # the whole refine STAGE (group + verify + decrypt + top-k) against the same roofline, and how much of it is record reuse:
    # reuse_factor = (query, candidate) pairs per distinct record decrypted in the batch
    refine_stage_ms = (stage["group"] + stage["verify"] + stage["decrypt"] + stage["topk"]) / K
    gpu.touched(clear=True)
    step_dev(Wm)
    barrier()
    distinct = int(gpu.touched(clear=True).size)
    pairs_one = int(d_cnt[:, 5].sum().item())
    reuse = pairs_one / max(distinct, 1)
    stage_achieved = alg_bytes / (refine_stage_ms * 1e-3) / 1e9 if refine_stage_ms > 0 else 0.0

    # ---- no_reuse: the same refine stage when NO record is shared between pairs -- the candidate lists are a permutation of the store
    #      (every record decrypted once and scored once), through fspann_refine_batch_dev.  This is what config 4's 100M records look like.
    no_reuse = None
    if rank == 0:
        Bn = cfg.B
        Qn = min(Q, cfg.N // Bn)
        if Qn >= 8:
            perm = np.random.default_rng(12345).permutation(cfg.N)[: Qn * Bn].astype(np.int32).reshape(Qn, Bn)
            d_cand = torch.from_numpy(perm).cuda()
            d_nc = torch.full((Qn,), Bn, dtype=torch.int32, device="cuda")
            d_rk = torch.empty((Qn, k), dtype=torch.int32, device="cuda")
            d_nd = torch.empty((Qn,), dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()

            def step_nr(i):
                gpu.refine_batch_dev(Qn, d_batches[i % n_batches].data_ptr(), d_cand.data_ptr(), d_nc.data_ptr(), Bn, k, d_ids.data_ptr(), d_dist.data_ptr(),
                                     d_rk.data_ptr(), d_nret.data_ptr(), d_nd.data_ptr())
            for i in range(Wm):
                step_nr(i)
            torch.cuda.synchronize(); gpu.sync()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0.record(stream)
            for i in range(K):
                step_nr(Wm + i)
            n1.record(stream)
            torch.cuda.synchronize(); gpu.sync()
            nr_ms = n0.elapsed_time(n1) / K
            nr_bytes = Qn * Bn * (8 * dim + 36) + Qn * (8 * dim + 12 * k)
            ok_nr = bool((d_nd == Bn).all().item())
            no_reuse = {"ms_per_launch": nr_ms, "pairs": Qn * Bn, "distinct_records": Qn * Bn, "algorithmic_bytes": nr_bytes,
                        "achieved": nr_bytes / (nr_ms * 1e-3) / 1e9, "frac": nr_bytes / (nr_ms * 1e-3) / 1e9 / peak, "all_decrypted": ok_nr,
                        "what": "refine stage (group + verify + decrypt + top-k) with candidate lists = a permutation of the store: reuse factor 1.0",
                        "limiter": "AES-256-GCM without AES/CLMUL instructions is bound by shared-memory table look-ups (LSU), not by HBM; see profiles/"}
            gpu.touched(clear=True)
            del d_cand

    # ---- e2e: host buffers through the public C-ABI call (pinned queries -> H2D -> search -> D2H results)
    h_q = [torch.from_numpy(b).pin_memory() for b in batches]
    # pinned pages that were never a DMA source copy slower the first time (and some allocations sit on slower host memory: tools/pin_test.py):
    # touch every buffer once, then record what one upload of each costs -- the timed region below still uploads every step's batch in full
    scratch = torch.empty_like(d_batches[0])
    h2d_ms = []
    for rep in range(2):
        h2d_ms = []
        for hb in h_q:
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(); scratch.copy_(hb, non_blocking=True); c1.record(); torch.cuda.synchronize()
            h2d_ms.append(round(c0.elapsed_time(c1), 3))
    del scratch
    h_ids = torch.empty((Q, k), dtype=torch.int32).pin_memory()
    h_dist = torch.empty((Q, k), dtype=torch.float64).pin_memory()
    h_nret = torch.empty((Q,), dtype=torch.int32).pin_memory()
    h_cnt = torch.empty((Q, 6), dtype=torch.int64).pin_memory()

    def step_host(i):
        gpu.search_batch_raw(Q, h_q[i % n_batches].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 0, h_ids.data_ptr(), h_dist.data_ptr(),
                             h_nret.data_ptr(), h_cnt.data_ptr())
    sampler.start()
    for i in range(Wm):
        step_host(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    f0.record(stream)
    step_wall = []
    for i in range(K):
        t_s = time.perf_counter()
        step_host(Wm + i)                                                 # returns synchronised: results are in the host buffers
        step_wall.append((time.perf_counter() - t_s) * 1e3)
    f1.record(stream)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    e2e_wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop()
    checksum = int(h_ids.to(torch.int64).sum().item())
    # the host-buffer call (chunked upload overlapped with TokenGen + Route) must return what the device-resident call returns
    step_dev(Wm + K - 1)
    barrier()
    e2e_same = bool(torch.equal(h_ids, d_ids.cpu()) and torch.equal(h_dist.view(torch.int64), d_dist.cpu().view(torch.int64)) and
                    torch.equal(h_nret, d_nret.cpu()) and torch.equal(h_cnt, d_cnt.cpu()))

    # ---- recall@10 of what was just timed: exact ground truth and recall on the device, outside the timed region
    recall = None
    if rank == 0:
        nq_gt = min(Q, 2000)
        t0 = time.time()
        gt = gpu.groundtruth(world["base"].astype(np.float32), batches[0][:nq_gt].astype(np.float32), k)
        res = gpu.search_batch(batches[0][:nq_gt], k, cfg.probes, cfg.hard_cap, cfg.B)
        rec = gpu.recall_batch(gt, res["top_ids"], k, res["n_ret"])
        recall = {"recall_at_k": float(rec.mean()), "k": k, "queries": nq_gt, "groundtruth_s": time.time() - t0,
                  "definition": "GroundtruthPrecompute.run + computeMetricsAtK (FSA:785-794), evaluated on the device"}
        log(f"recall@{k} = {recall['recall_at_k']:.4f} over {nq_gt} queries (ground truth + search + recall in {recall['groundtruth_s']:.1f}s)")


    # ---- strong scaling of ONE batch: (a) split N ways over replicated state (config 3's split, no collective);
    #      (b) the store sharded by id range across the N GPUs through fspann_sharded_search_batch_dev (config 4's split: NCCL
    #      all-gathers of the candidate lists and of the per-shard top-k INSIDE the C library).  Same batches on every rank.
    sharded = strong = None
    if not args.no_sharded:
        from fspann_query_system_b200 import distributed as DD
        common = batches if rank == 0 else make_batches(cfg, n_batches, 0)
        d_common = d_batches if rank == 0 else [torch.from_numpy(b).cuda() for b in common]
        run = DevRunner(gpu, torch, local_rank, Q, k)
        refs = []
        for b in range(min(2, n_batches)):                               # unsharded results of the same batches (replicated store, every rank)
            gpu.search_batch_dev(Q, d_common[b].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 0, 1, run.ids.data_ptr(), run.dist.data_ptr(),
                                 run.nret.data_ptr(), run.cnt.data_ptr())
            refs.append(run.result())
        if world_size > 1:
            qlo, qhi = DD.split_batch(Q, rank, world_size)

            def step_sr(i):
                gpu.search_batch_dev(qhi - qlo, d_common[i % n_batches].data_ptr() + qlo * dim * 8, k, cfg.probes, cfg.hard_cap, cfg.B, 0, 1,
                                     run.ids.data_ptr(), run.dist.data_ptr(), run.nret.data_ptr(), run.cnt.data_ptr())
            sr_ms = run.timed(step_sr, K, Wm, barrier)
            t = torch.tensor([sr_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sr_ms = float(t.item())
            strong = {"ms_per_batch": sr_ms / K, "queries_per_s": Q * K / (sr_ms * 1e-3), "queries_per_gpu": qhi - qlo,
                      "what": f"one {Q}-query batch split {world_size} ways, routing index + store replicated, no collective (max over ranks)"}
        slo, shi = DD.shard_range(cfg.N, rank, world_size)
        gpu.store_upload(dim, world["iv"][slo:shi], world["ct"][slo:shi], world["ver"][slo:shi], id_base=slo, n_global=cfg.N)
        cid = [gpu.comm_unique_id() if (rank == 0 and world_size > 1) else None]
        if world_size > 1:
            dist.broadcast_object_list(cid, src=0)
        gpu.comm_init(world_size, rank, cid[0])

        def step_sh(i):
            gpu.sharded_search_batch_dev(Q, d_common[i % n_batches].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 1, run.ids.data_ptr(),
                                         run.dist.data_ptr(), run.nret.data_ptr(), run.cnt.data_ptr())
        sh_l0 = gpu.launch_count()
        sh_ms = run.timed(step_sh, K, Wm, barrier)
        sh_launches = (gpu.launch_count() - sh_l0) / (K + Wm)
        # every call ends in a host synchronisation (the retry decision), so one rank's host hiccup stalls all ranks in the next all-gather:
        # besides the mean over the K steps, time a few steps one by one and keep the median and the stage split of the median step
        per_step = []
        for i in range(max(K, 9)):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(run.stream)
            step_sh(i)
            e1.record(run.stream)
            barrier()
            st = gpu.sharded_stage_ms()
            st["refine_split"] = {kk: vv for kk, vv in gpu.stage_ms().items() if kk in ("group", "verify", "decrypt", "topk")}
            per_step.append((e0.elapsed_time(e1), st))
        per_step.sort(key=lambda t: t[0])
        sh_med, sh_stage = per_step[len(per_step) // 2]
        sh_refine = sh_stage.pop("refine_split")
        eq = True
        for b in range(len(refs)):
            step_sh(b)
            eq = eq and same_result(run.result(), refs[b])
        if world_size > 1:
            t = torch.tensor([sh_ms, 0.0 if eq else 1.0, sh_med], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sh_ms, eq, sh_med = float(t[0].item()), bool(t[1].item() == 0.0), float(t[2].item())
        sharded = {"n_shards": world_size, "ms_per_batch": sh_ms / K, "ms_per_batch_median": sh_med, "queries_per_s": Q * K / (sh_ms * 1e-3), "equals_unsharded": eq,
                   "stage_ms": {"route_my_slice": sh_stage["route"], "allgather_candidates": sh_stage["allgather_candidates"],
                                "refine_my_shard": sh_stage["refine"], "allgather_topk_and_merge": sh_stage["allgather_topk_merge"], "refine_split": sh_refine},
                   "allgather_bytes": sh_stage["gather_bytes"], "allgather_ms": sh_stage["allgather_candidates"] + sh_stage["allgather_topk_merge"],
                   "gpu_launches_per_batch": sh_launches, "records_per_shard": shi - slo,
                   "what": f"one {Q}-query batch, store sharded by id range over {world_size} GPU(s), routing index replicated; "
                           "fspann_sharded_search_batch_dev: NCCL all-gather x2 + all-reduce inside libfspann_gpu.so (max over ranks)"}
        log(f"sharded x{world_size}: {sharded['ms_per_batch']:.3f} ms/batch (median of single steps {sh_med:.3f}), equals unsharded: {eq}; stages {sharded['stage_ms']}")
        gpu.comm_destroy()
        if world_size == 1:                                               # later legs (config 5, CPU sample) use the whole store again
            gpu.store_upload(dim, world["iv"], world["ct"], world["ver"])

    # ---- config 5: Rotate -> v2 + partial Migrate on the device, then the same batch again: results must not move (routing-ciphertext
    #      orthogonality) and the per-record-version path must cost the same.  Outside the timed regions above.
    mixed = None
    if rank == 0 and world_size == 1 and not args.no_mixed:
        ref_ids = gpu.search_batch(batches[0], k, cfg.probes, cfg.hard_cap, cfg.B)
        km = HS.KeyManager(WL.MASTER_KEY)
        gpu.keys_set(2, km.derive(2))                                                   # Rotate: v2 becomes available
        mig = np.arange(0, cfg.N, 3, dtype=np.int32)
        t0 = time.time()
        out = gpu.migrate(mig, WL.record_ivs(len(mig), cfg.base_seed + 77), 2)         # Migrate ids = 0 (mod 3), in place in HBM
        t_mig = time.time() - t0
        for i in range(Wm):
            step_dev(i)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for i in range(K):
            step_dev(Wm + i)
        g1.record(stream)
        barrier()
        again = gpu.search_batch(batches[0], k, cfg.probes, cfg.hard_cap, cfg.B)
        same = bool(np.array_equal(again["top_ids"], ref_ids["top_ids"]) and np.array_equal(again["top_dist"].view(np.uint64), ref_ids["top_dist"].view(np.uint64))
                    and np.array_equal(again["counters"], ref_ids["counters"]))
        mixed = {"migrated_records": int(out["count"]), "migrate_s_incl_copies": t_mig, "results_identical_after_rotate_migrate": same,
                 "ms_per_step_mixed_versions": g0.elapsed_time(g1) / K, "versions": [1, 2]}
        step_dev(0)
        mixed["stage_ms"] = {kk: vv for kk, vv in gpu.stage_ms().items() if kk != "launches"}
        log(f"config 5: migrated {out['count']} records to v2 on the device in {t_mig:.2f}s; results identical: {same}; {mixed['ms_per_step_mixed_versions']:.2f} ms/step")

    if world_size > 1:
        t = torch.tensor([ms, e2e_ms, e2e_wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, e2e_wall_ms = (float(x) for x in t.tolist())
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln.item())
    total_q = Q * K * world_size
    value = total_q / (ms * 1e-3)
    e2e_val = total_q / (e2e_ms * 1e-3)

    # ---- CPU baseline beside it: the oracle port on this box's host cores (rank 0, N=1 only; bounded sample)
    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        O, ix, st = oracle_objects(cfg, world)
        threads = os.cpu_count() or 1
        nq = min(Q, max(64, threads * 96))
        dt, ref = cpu_run(cfg, O, ix, st, batches[0][:nq], threads)
        dt1, _ = cpu_run(cfg, O, ix, st, batches[0][:min(nq, 96)], 1)
        # the GPU result for the same queries must be identical (ids exact, FP64 distances bit-exact)
        got = gpu.search_batch(batches[0][:nq], k, cfg.probes, cfg.hard_cap, cfg.B)
        same = bool(np.array_equal(got["top_ids"], ref["top_ids"]) and np.array_equal(got["top_dist"].view(np.uint64), ref["top_dist"].view(np.uint64)))
        cpu_baseline = {"value": nq / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                        "sample": f"first {nq} queries of batch 0, {threads} threads; single-thread: {min(nq, 96) / dt1:.1f} queries/s on {min(nq, 96)} queries",
                        "single_thread_value": min(nq, 96) / dt1, "gpu_matches_oracle_on_sample": same}


    # ---- the other BASELINE configs and the reference's published operating point, one GPU each (N=1 runs only)
    configs_out = None
    if rank == 0 and world_size == 1 and not args.no_extras:
        configs_out = {}
        threads = os.cpu_count() or 1
        for name in [x for x in args.extras.split(",") if x]:
            try:
                configs_out[name] = run_extra_config(name, WL.CONFIGS[name], torch, local_rank, 3, 2, threads, log)
            except Exception as e:      # noqa: BLE001  -- an extra leg must never take the headline number down with it
                configs_out[name] = {"error": f"{type(e).__name__}: {e}"}
                log(f"config {name} failed: {e}")
        if "P6" in configs_out and "error" not in configs_out["P6"]:
            configs_out["P6"]["published"] = {"art_ms_per_query": 2828.0, "recall_at_10": 0.838, "hardware": "Xeon E5-2630 v4, Java 21, one query at a time",
                                              "source": "README.md:300, config_sift1m.json:59-71", "note": "published, different hardware and real SIFT1M data"}

    if rank == 0:
        out = {"metric": metric, "value": value, "unit": "queries/s", "n_gpus": world_size, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "clocks": clocks, "gpu_launches": launches,
               "e2e": {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": Q * dim * 8, "d2h_bytes_per_step": Q * k * 12 + Q * 4 + Q * 48,
                       "ms_per_step": e2e_ms / K, "wall_ms_per_step": e2e_wall_ms / K, "wall_ms_of_each_step": [round(v, 3) for v in step_wall], "h2d_ms_of_each_pinned_batch": h2d_ms, "result_checksum": checksum,
                       "equals_device_path": e2e_same},
               "roofline": {"bound": "hbm", "kernel": "refine_decrypt_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak if peak else None, "traffic": load_traffic(args.config == "C2" and not args.n and not args.q),
                            "traffic_unit": "bytes/launch (dram read+write, ncu; profiles/r2_decrypt_traffic.json)", "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dec_ms, "pairs_per_launch": pairs / K,
                            "stage_ms": refine_stage_ms, "stage_achieved": stage_achieved, "stage_frac": stage_achieved / peak if peak else None,
                            "stage_what": "group + verify + decrypt + top-k: the whole refine stage, same algorithmic bytes",
                            "reuse_factor": reuse, "distinct_records_per_launch": distinct,
                            "reuse_note": "records are decrypted once per batch and scored once per (query, candidate) pair; frac and stage_frac count the "
                                          "algorithmic bytes of every pair, so they scale with reuse_factor -- no_reuse is the same stage at reuse 1.0",
                            "no_reuse": no_reuse},
               "stage_ms_per_step": {s: v / K for s, v in stage.items()},
               "recall": recall, "setup_on_device_s": world["setup"], "mixed_versions": mixed,
               "sharded": sharded, "strong_replicated": strong, "configs": configs_out,
               "cpu_baseline": cpu_baseline}
        print(json.dumps(out), file=real_stdout, flush=True)
    gpu.close()
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
