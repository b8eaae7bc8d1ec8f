"""Committed golden vectors (tests/golden/fspann_small.npz, made by tests/golden/make_golden.py from the oracle; the
reference holds none for this path).  CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fspann_small.npz"))
N, dim, Q, T, D, m, lam, k, B, probes, hard_cap = (int(x) for x in G["params"])


def objects():
    g = O.GFunctions(dim, T, D, m, lam, G["alpha"], G["r"], G["omega"])
    ix = O.Index(g, N, G["min_key"].shape[1], G["min_key"], G["max_key"], G["rep"], G["ids"])
    st = O.Store(dim, G["iv"], G["ct"], G["key_version"], {1: G["key1"].tobytes(), 2: G["key2"].tobytes()})
    return g, ix, st


def test_oracle_reproduces_golden():
    g, ix, st = objects()
    assert np.array_equal(O.tokengen_batch(G["queries"], g), G["qcodes"])
    assert np.array_equal(O.tokengen_batch(G["base"], g), G["base_codes"])
    ix2 = O.index_build(G["base_codes"], g, O.staged_order(N))
    assert np.array_equal(ix2.ids, G["ids"]) and np.array_equal(ix2.rep, G["rep"]) and np.array_equal(ix2.min_key, G["min_key"])
    for q in range(Q):
        ids, sc, raw, _ = O.route(ix, G["qcodes"][q], probes, hard_cap)
        n = G["ncand"][q]
        assert np.array_equal(ids[:n], G["cand"][q, :n]) and np.array_equal(sc[:n], G["cand_scores"][q, :n])
        assert raw == G["raw"][q] and len(ids) == G["uniq"][q]
        s = O.search(ix, st, G["queries"][q], G["qcodes"][q], k, probes, hard_cap, B)
        assert np.array_equal(s["verdict"], G["verdict"][q, :n])
        assert np.array_equal(s["top_ids"], G["top_ids"][q, :G["nret"][q]])
        assert np.array_equal(s["top_dist"].view(np.uint64), G["top_dist"][q, :G["nret"][q]].view(np.uint64))
    assert (G["verdict"] == O.VERDICT_TAG_FAIL).sum() >= 1


@pytest.mark.gpu
def test_cuda_path_reproduces_golden():
    from fspann_query_system_b200.gpu import GpuContext
    ctx = GpuContext(0)
    try:
        ctx.routing_upload(dim, T, D, m, lam, G["alpha"], G["r"], G["omega"], G["min_key"], G["max_key"], G["rep"], G["ids"])
        ctx.keys_set(1, G["key1"].tobytes()); ctx.keys_set(2, G["key2"].tobytes())
        ctx.store_upload(dim, G["iv"], G["ct"], G["key_version"])
        assert np.array_equal(ctx.tokengen_batch(G["queries"]), G["qcodes"])
        assert np.array_equal(ctx.tokengen_batch(G["base"]), G["base_codes"])
        for general in (0, 1):
            ctx.set_option("route_general", general)
            r = ctx.route_batch(G["qcodes"], probes, hard_cap, B)
            assert np.array_equal(r["n_cand"], G["ncand"]) and np.array_equal(r["raw_seen"], G["raw"]) and np.array_equal(r["unique"], G["uniq"])
            for q in range(Q):
                n = G["ncand"][q]
                assert np.array_equal(r["cand_ids"][q, :n], G["cand"][q, :n]) and np.array_equal(r["cand_scores"][q, :n], G["cand_scores"][q, :n])
        ctx.set_option("route_general", 0)
        f = ctx.refine_batch(G["queries"], G["cand"], G["ncand"], k)
        s = ctx.search_batch(G["queries"], k, probes, hard_cap, B)
        for out in (f, s):
            assert np.array_equal(out["n_ret"], G["nret"])
            for q in range(Q):
                nr = G["nret"][q]
                assert np.array_equal(out["top_ids"][q, :nr], G["top_ids"][q, :nr])
                assert np.array_equal(out["top_dist"][q, :nr].view(np.uint64), G["top_dist"][q, :nr].view(np.uint64))
        for q in range(Q):
            assert np.array_equal(f["verdict"][q, :G["ncand"][q]], G["verdict"][q, :G["ncand"][q]])
    finally:
        ctx.close()


# ---- SURVEY 8(f) additions: Migrate, ground truth, recall (tests/golden/fspann_8f.npz, made by make_golden_8f.py)
F8 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fspann_8f.npz"))


def test_oracle_reproduces_golden_8f():
    keys = {1: G["key1"].tobytes(), 2: G["key2"].tobytes(), 3: F8["key3"].tobytes()}
    st = O.Store(dim, G["iv"].copy(), G["ct"].copy(), G["key_version"].copy(), keys)
    ids = F8["migrate_ids"]
    assert O.migrate(st, np.where((ids >= 0) & (ids < N), ids, -1), F8["fresh_ivs"], 3) == int(F8["migrated"])
    assert np.array_equal(st.iv, F8["iv_after"]) and np.array_equal(st.ct, F8["ct_after"]) and np.array_equal(st.key_version, F8["ver_after"])
    gt, d2 = O.groundtruth(G["base"].astype(np.float32), G["queries"].astype(np.float32), 8)
    assert np.array_equal(gt, F8["gt_ids"]) and np.array_equal(d2.view(np.uint64), F8["gt_d2"].view(np.uint64))
    rec = [O.recall_at_k(gt[q, :5], G["top_ids"][q], int(G["nret"][q]), 5) for q in range(Q)]
    assert np.array_equal(np.array(rec), F8["recall5"])


@pytest.mark.gpu
def test_cuda_path_reproduces_golden_8f():
    from fspann_query_system_b200.gpu import GpuContext
    ctx = GpuContext(0)
    try:
        ctx.routing_upload(dim, T, D, m, lam, G["alpha"], G["r"], G["omega"], G["min_key"], G["max_key"], G["rep"], G["ids"])
        for v, kk in ((1, G["key1"]), (2, G["key2"]), (3, F8["key3"])):
            ctx.keys_set(v, kk.tobytes())
        ctx.store_upload(dim, G["iv"], G["ct"], G["key_version"])
        out = ctx.migrate(F8["migrate_ids"], F8["fresh_ivs"], 3)
        assert out["count"] == int(F8["migrated"])
        sel = np.nonzero(out["reencrypted"])[0]
        rows = F8["migrate_ids"][sel]
        assert np.array_equal(out["iv"][sel], F8["iv_after"][rows]) and np.array_equal(out["ct"][sel], F8["ct_after"][rows])
        assert (F8["ver_after"][rows] == 3).all() and len(sel) == int((F8["ver_after"] != G["key_version"]).sum())
        s = ctx.search_batch(G["queries"], k, probes, hard_cap, B)               # routing-ciphertext orthogonality: the golden results hold
        for q in range(Q):
            nr = G["nret"][q]
            assert s["n_ret"][q] == nr and np.array_equal(s["top_ids"][q, :nr], G["top_ids"][q, :nr])
        gt, d2 = ctx.groundtruth(G["base"].astype(np.float32), G["queries"].astype(np.float32), 8, want_d2=True)
        assert np.array_equal(gt, F8["gt_ids"]) and np.array_equal(d2.view(np.uint64), F8["gt_d2"].view(np.uint64))
        rec = ctx.recall_batch(gt[:, :5].copy(), G["top_ids"], 5, G["nret"])
        assert np.array_equal(rec, F8["recall5"])
    finally:
        ctx.close()
