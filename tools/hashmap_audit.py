#!/usr/bin/env python
"""java.util.HashMap chain-length audit for decimal-string ids (hard part 2).

The reference's candidate order depends on HashMap<String,...> iteration order (PIS:413-420 staging map, PIS:619 bestScore map).  A bin
is TREEIFIED by the JDK when its 9th entry arrives (putVal: binCount >= TREEIFY_THRESHOLD - 1 with 8 nodes already chained, table >= 64);
treeify() then moves the red-black root to the FRONT of the bin's next-list (moveRootToFront), so the bin no longer iterates in
insertion order -- which neither the oracle nor the kernels model.  This script computes, for ids "0".."N-1", the longest chain of the staging map new HashMap<>(N)
(cap = tableSizeFor(N), doubled while N > 0.75*cap) -- pure numpy, no oracle, no CUDA.

  python tools/hashmap_audit.py 100000000
"""
import sys

import numpy as np


def java_hash_decimal(ids: np.ndarray) -> np.ndarray:
    """String.hashCode of the decimal string of each non-negative id, then HashMap.hash's spread h ^ (h >>> 16); uint32."""
    u = ids.astype(np.uint64)
    h = np.zeros(ids.shape, dtype=np.uint64)
    p = np.ones(ids.shape, dtype=np.uint64)
    alive = np.ones(ids.shape, dtype=bool)
    M = np.uint64(0xFFFFFFFF)
    while alive.any():
        digit = u % np.uint64(10)
        h = np.where(alive, (h + (np.uint64(48) + digit) * p) & M, h)
        p = np.where(alive, (p * np.uint64(31)) & M, p)
        u = u // np.uint64(10)
        alive &= u > 0
    h32 = h.astype(np.uint32)
    return h32 ^ (h32 >> np.uint32(16))


def table_size_for(n: int) -> int:
    c = 1
    while c < n:
        c <<= 1
    return c


def audit(N: int, chunk: int = 10_000_000):
    cap = table_size_for(N)
    while N > 0.75 * cap:
        cap <<= 1
    counts = np.zeros(cap, dtype=np.uint8)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        b = java_hash_decimal(np.arange(lo, hi, dtype=np.int64)) & np.uint32(cap - 1)
        counts += np.bincount(b, minlength=cap).astype(np.uint8)
    hist = np.bincount(counts)
    return cap, int(counts.max()), hist, np.nonzero(counts >= 9)[0]


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    cap, mx, hist, big = audit(N)
    print(f"N={N} cap=2^{cap.bit_length() - 1} max_chain={mx} treeified_bins(>=9)={int(hist[9:].sum())} chain_histogram={hist.tolist()} treeified_bucket_indices={big.tolist()[:16]}")
