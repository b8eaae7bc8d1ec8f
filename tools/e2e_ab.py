"""A/B of the host-buffer search entry with and without the overlapped (chunked) query upload.  python tools/e2e_ab.py [C2]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from fspann_query_system_b200 import workloads as WL  # noqa: E402
from fspann_query_system_b200.gpu import GpuContext  # noqa: E402


def main():
    cfg = getattr(WL, sys.argv[1] if len(sys.argv) > 1 else "C2")
    gpu = GpuContext(0)
    world, batches = bench.build_world(cfg, gpu, 3, 0, quiet=True)
    Q, k = batches[0].shape[0], cfg.k
    h_q = [torch.from_numpy(b).pin_memory() for b in batches]
    h_ids = torch.empty((Q, k), dtype=torch.int32).pin_memory()
    h_dist = torch.empty((Q, k), dtype=torch.float64).pin_memory()
    h_nret = torch.empty((Q,), dtype=torch.int32).pin_memory()
    h_cnt = torch.empty((Q, 6), dtype=torch.int64).pin_memory()

    def step(i):
        gpu.search_batch_raw(Q, h_q[i % 3].data_ptr(), k, cfg.probes, cfg.hard_cap, cfg.B, 0, h_ids.data_ptr(), h_dist.data_ptr(), h_nret.data_ptr(),
                             h_cnt.data_ptr())

    for mode in (0, 2, 3, 4, 0, 2):
        gpu.set_option("h2d_overlap", mode)
        for i in range(3):
            step(i)
        t0 = time.perf_counter()
        for i in range(10):
            step(i)
        dt = (time.perf_counter() - t0) / 10
        print(f"h2d_overlap={mode}: {dt * 1e3:.3f} ms per call, {Q / dt / 1e6:.3f} M q/s, checksum {int(h_ids.to(torch.int64).sum())}", flush=True)


if __name__ == "__main__":
    main()
