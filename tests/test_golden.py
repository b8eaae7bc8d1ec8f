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
