"""Generates tests/golden/fspann_small.npz: inputs and expected outputs of the hot path for one small deployment.

The reference is pure Java and cannot run in this image (no JVM), and its own tests hold no golden vectors for this path
(SURVEY.md 8c), so these vectors come from the ORACLE (oracle/fspann_oracle.c) -- "parity unpinned" by the reference,
pinned only as a regression fixture: the oracle (CPU test) and the CUDA path (GPU test) must both reproduce them.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    rng = np.random.Generator(np.random.PCG64(20261018))
    N, dim, Q, T, D, m, lam, k, B, probes, hard_cap = 1200, 16, 12, 2, 3, 12, 2, 5, 64, 5, 20000
    cen = rng.uniform(0, 128, size=(16, dim)).astype(np.float32)
    base = np.rint(np.clip(cen[rng.integers(0, 16, N)] + rng.normal(0, 12, (N, dim)).astype(np.float32), 0, 255)).astype(np.float64)
    queries = np.rint(np.clip(cen[rng.integers(0, 16, Q)] + rng.normal(0, 12, (Q, dim)).astype(np.float32), 0, 255)).astype(np.float64)
    g = O.registry_init(base[:1000], m, lam, 13, T, D)
    codes = O.tokengen_batch(base, g)
    ix = O.index_build(codes, g, O.staged_order(N))
    master = bytes(range(32))
    keys = {1: O.kdf(master, 1), 2: O.kdf(master, 2)}
    iv = rng.integers(0, 256, size=(N, 12), dtype=np.uint8)
    ver = np.where(np.arange(N) % 3 == 0, 2, 1).astype(np.int32)
    ct = np.empty((N, 8 * dim + 16), dtype=np.uint8)
    for v in keys:
        sel = np.nonzero(ver == v)[0].astype(np.int32)
        ct[sel] = O.encrypt_store(base[sel], v, keys[v], iv[sel], ids=sel)
    qcodes = O.tokengen_batch(queries, g)
    victim = int(O.route(ix, qcodes[0], probes, hard_cap)[0][3])
    ct[victim, 5] ^= 1                  # one tampered record among query 0's candidates: must come back as a tag failure
    store = O.Store(dim, iv, ct, ver, keys)
    cand = np.full((Q, B), -1, np.int32); csc = np.full((Q, B), -1, np.int32); ncand = np.zeros(Q, np.int32)
    raw = np.zeros(Q, np.int32); uniq = np.zeros(Q, np.int32); verdict = np.full((Q, B), 255, np.uint8)
    top_ids = np.full((Q, k), -1, np.int32); top_dist = np.full((Q, k), np.nan); nret = np.zeros(Q, np.int32)
    for q in range(Q):
        ids, sc, r, _ = O.route(ix, qcodes[q], probes, hard_cap)
        n = min(B, len(ids))
        cand[q, :n], csc[q, :n], ncand[q], raw[q], uniq[q] = ids[:n], sc[:n], n, r, len(ids)
        s = O.search(ix, store, queries[q], qcodes[q], k, probes, hard_cap, B)
        assert not s["retried"]
        verdict[q, :n] = s["verdict"]
        nret[q] = len(s["top_ids"]); top_ids[q, :nret[q]] = s["top_ids"]; top_dist[q, :nret[q]] = s["top_dist"]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fspann_small.npz")
    np.savez_compressed(out, params=np.array([N, dim, Q, T, D, m, lam, k, B, probes, hard_cap]), base=base, queries=queries, alpha=g.alpha, r=g.r,
                        omega=g.omega, min_key=ix.min_key, max_key=ix.max_key, rep=ix.rep, ids=ix.ids, iv=iv, ct=ct, key_version=ver,
                        key1=np.frombuffer(keys[1], np.uint8), key2=np.frombuffer(keys[2], np.uint8), base_codes=codes, qcodes=qcodes, cand=cand,
                        cand_scores=csc, ncand=ncand, raw=raw, uniq=uniq, verdict=verdict, top_ids=top_ids, top_dist=top_dist, nret=nret)
    print("wrote", out, os.path.getsize(out), "bytes; tag failures among refined candidates:", int((verdict == 3).sum()))


if __name__ == "__main__":
    main()
