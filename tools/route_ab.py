"""A/B of the two fast Route kernels on the bench workload (C2 by default): stage times of a search step with the one-CTA kernel
(route_v1 = 1) and the two-CTA kernel, and how many queries the latter handed back.  Run on a GPU box: python tools/route_ab.py [C2|C3]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from fspann_query_system_b200 import workloads as WL  # noqa: E402
from fspann_query_system_b200.gpu import GpuContext  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    cfg = getattr(WL, name)
    gpu = GpuContext(0)
    for a in sys.argv[2:]:                                     # further library options: name=value
        k_, v_ = a.split("=")
        gpu.set_option(k_, int(v_))
    world, batches = bench.build_world(cfg, gpu, 3, 0, quiet=True)
    Q, k = batches[0].shape[0], cfg.k
    run = bench.DevRunner(gpu, torch, 0, Q, k)
    d_batches = [torch.from_numpy(b).cuda() for b in batches]

    def step(i):
        gpu.search_batch_dev(Q, d_batches[i % 3].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 0, 1, run.ids.data_ptr(), run.dist.data_ptr(),
                             run.nret.data_ptr(), run.cnt.data_ptr())

    def barrier():
        torch.cuda.synchronize(); gpu.sync()

    res = {}
    for v1 in (1, 0, 1, 0):
        gpu.set_option("route_v1", v1)
        ms = run.timed(step, 6, 3, barrier) / 6
        stage = gpu.stage_ms()
        print(f"route_v1={v1}: {ms:.3f} ms/step, stages { {kk: round(vv, 3) for kk, vv in stage.items()} }, v2={gpu.get_info('last_route_v2')} overflowed={gpu.get_info('route_overflowed')}", flush=True)
        step(0)
        res[v1] = run.result()
    print("identical:", bench.same_result(res[0], res[1]))


if __name__ == "__main__":
    main()
