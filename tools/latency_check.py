#!/usr/bin/env python
"""Small-batch latency of fspann_search_batch (host buffers in, host buffers out) at BASELINE config 2: the reference's own metric is
per-query latency (ART, README), so this reports ms per call for Q = 1, 8, 64, 512 next to the 10k-batch throughput of bench.py."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fspann_query_system_b200 import hostsetup as HS, workloads as WL  # noqa: E402
from fspann_query_system_b200.gpu import GpuContext  # noqa: E402

cfg = WL.CONFIGS[os.environ.get("FSPANN_BENCH_CONFIG", "C2")]
base = WL.base_vectors(cfg)
alpha, r, omega = HS.build_gfunctions(base[:HS.MIN_SAMPLE_SIZE], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
gpu = GpuContext(0)
gpu.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
gpu.routing_build(base, HS.staged_order(cfg.N), want_arrays=False)
km = HS.KeyManager(WL.MASTER_KEY)
gpu.keys_set(1, km.derive(1))
iv = WL.record_ivs(cfg.N, cfg.base_seed + 5)
gpu.store_upload(cfg.dim, iv, gpu.encrypt_batch(np.arange(cfg.N, dtype=np.int32), base, iv, 1), np.ones(cfg.N, dtype=np.int32))
queries = WL.query_vectors(cfg)
out = {}
for Q in (1, 8, 64, 512, 4096):
    for i in range(5):
        gpu.search_batch(queries[i * Q:(i + 1) * Q], cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
    n = 20
    t0 = time.perf_counter()
    for i in range(n):
        s = (i * Q) % (cfg.Q - Q + 1)
        gpu.search_batch(queries[s:s + Q], cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
    dt = (time.perf_counter() - t0) / n
    out[f"Q={Q}"] = {"ms_per_call": 1e3 * dt, "queries_per_s": Q / dt}
print(json.dumps({"workload": cfg.name, "latency": out}))
gpu.close()
