#!/usr/bin/env python
"""Tensor-core TokenGen (tcgen05) vs the exact FP64 kernel: codes must be identical.  python tools/tc_check.py [N]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fspann_query_system_b200 import hostsetup as HS, workloads as WL
from fspann_query_system_b200.gpu import GpuContext

def run(cfg, n):
    base = WL.base_vectors(cfg.scaled(N=max(n, 1000)))[:n]
    alpha, r, omega = HS.build_gfunctions(base[:1000], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    g = GpuContext(0)
    g.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
    out = {}
    for mode, name in ((1, "exact"), (2, "fp32"), (0, "tensor")):
        g.set_option("tokengen_mode", mode)
        g.tokengen_batch(base[:256])
        t0 = time.time(); c = g.tokengen_batch(base); dt = time.time() - t0
        out[name] = c
        print(f"{cfg.name} n={n} mode={name} path={g.get_info('last_tokengen_path')} rechecked={g.get_info('tokengen_rechecked')} "
              f"overflow={g.get_info('tokengen_overflow')} wall={dt*1e3:.1f} ms", flush=True)
    ok = np.array_equal(out["tensor"], out["exact"]) and np.array_equal(out["fp32"], out["exact"])
    print("IDENTICAL" if ok else f"MISMATCH: {int((out['tensor'] != out['exact']).sum())} code words differ", flush=True)
    g.close()
    return ok

def timed(cfg, n):
    """Device time of TokenGen over n vectors already resident in HBM (CUDA events on the library stream), per mode."""
    import torch
    base = WL.base_vectors(cfg.scaled(N=max(n, 1000)))[:n]
    alpha, r, omega = HS.build_gfunctions(base[:1000], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    g = GpuContext(0)
    g.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
    dq = torch.from_numpy(base).cuda()
    dc = torch.zeros((n, cfg.T * cfg.D), dtype=torch.int64, device="cuda")
    stream = torch.cuda.ExternalStream(g.stream())
    torch.cuda.synchronize()
    res = {}
    for mode, name in ((1, "exact"), (2, "fp32"), (0, "tensor")):
        g.set_option("tokengen_mode", mode)
        for _ in range(2):
            g.tokengen_batch_dev(n, dq.data_ptr(), dc.data_ptr())
        g.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            g.tokengen_batch_dev(n, dq.data_ptr(), dc.data_ptr())
        e1.record(stream)
        g.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flops = 2.0 * n * cfg.T * cfg.D * cfg.m * cfg.dim
        res[name] = dict(ms=ms, algorithmic_tflops=flops / ms / 1e9, rechecked=g.get_info("tokengen_rechecked"))
        print(f"{cfg.name} n={n} {name}: {ms:.3f} ms  ({flops / ms / 1e9:.1f} algorithmic TFLOP/s), rechecked {res[name]['rechecked']}", flush=True)
    g.close()
    return res


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "time":
        import json
        n = int(sys.argv[1])
        print(json.dumps({"C2_queries_10k": timed(WL.C2, 10000), "C2_base_" + str(n): timed(WL.C2, n), "C3_base_" + str(n): timed(WL.C3, n)}))
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    ok = run(WL.C1, min(n, 10000)) & run(WL.C2, n) & run(WL.C3, n) & run(WL.C4S, n)
    sys.exit(0 if ok else 1)
