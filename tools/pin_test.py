"""Diagnostic: host-to-device copy time per pinned buffer (10 MB query batches).  Separately pinned tensors can land on slow memory (remote
NUMA node): bench.py's host-buffer leg therefore pins ONE block for all its batches and reports the copy time of every slice it uses."""
import numpy as np
import torch

d = torch.empty((10000, 128), dtype=torch.float64, device="cuda")


def times(bufs):
    ts = []
    for b in bufs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(b, non_blocking=True); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    return ts


sep = [torch.from_numpy(np.random.rand(10000, 128)).pin_memory() for _ in range(11)]
big = torch.empty((16, 10000, 128), dtype=torch.float64).pin_memory()
big.copy_(torch.from_numpy(np.random.rand(16, 10000, 128)))
torch.cuda.synchronize()
for rep in range(2):
    print("separately pinned:", times(sep))
    print("one pinned block :", times([big[i] for i in range(16)]))
