"""Generates tests/golden/fspann_8f.npz: expected outputs of the SURVEY 8(f) additions on the deployment of fspann_small.npz --
Migrate (reencryptTouched to v3 with fixed fresh IVs), exact ground truth (K = 8) and recall@5 of the golden search results.
Like fspann_small.npz these come from the ORACLE (no JVM here; the reference has no fixtures): a regression pin for both the oracle
(CPU test) and the CUDA path (GPU test).  Run from the repo root:  python tests/golden/make_golden_8f.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    G = np.load(os.path.join(here, "fspann_small.npz"))
    N, dim, Q, T, D, m, lam, k, B, probes, hard_cap = (int(x) for x in G["params"])
    rng = np.random.Generator(np.random.PCG64(8))
    master = bytes(range(32))
    keys = {1: G["key1"].tobytes(), 2: G["key2"].tobytes(), 3: O.kdf(master, 3)}
    st = O.Store(dim, G["iv"].copy(), G["ct"].copy(), G["key_version"].copy(), keys)
    ids = np.concatenate([rng.permutation(N)[:400], [5, 5, -1, N + 3]]).astype(np.int32)     # duplicates and foreign ids included
    fresh = rng.integers(0, 256, size=(len(ids), 12), dtype=np.uint8)
    done = O.migrate(st, np.where((ids >= 0) & (ids < N), ids, -1), fresh, 3)
    gt, d2 = O.groundtruth(G["base"].astype(np.float32), G["queries"].astype(np.float32), 8)
    rec = np.array([O.recall_at_k(gt[q, :5], G["top_ids"][q], int(G["nret"][q]), 5) for q in range(Q)])
    out = os.path.join(here, "fspann_8f.npz")
    np.savez_compressed(out, migrate_ids=ids, fresh_ivs=fresh, key3=np.frombuffer(keys[3], np.uint8), migrated=np.int64(done), iv_after=st.iv, ct_after=st.ct,
                        ver_after=st.key_version, gt_ids=gt, gt_d2=d2, recall5=rec)
    print("wrote", out, os.path.getsize(out), "bytes; migrated", done, "mean recall@5", rec.mean())


if __name__ == "__main__":
    main()
