"""GPU behaviour tests modelled on the reference's ITs: per-record key versions, Rotate / Migrate / Retire invariance
(ForwardSecurityGameTest G1-G5, ForwardSecurityAdversarialIT, config 5), per-candidate failure verdicts (QSI:242-270),
deleted points (PIS:739, PIS:718), error behaviour (SuperFailureModeIT) and the debug plaintext tap."""
import numpy as np
import pytest

from fspann_query_system_b200 import _native as N, hostsetup as HS, workloads as WL
from fspann_query_system_b200.api import (ForwardSecureANNSystem, PaperConfig, RuntimeConfig, SystemConfig)
from fspann_query_system_b200.gpu import GpuContext
from oracle import oracle as O

pytestmark = pytest.mark.gpu

W3 = dict(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2, n_versions=3)
W1 = dict(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2)


def check_search(ctx, w, store, k=10, B=64, probes=5, hard_cap=20000, deleted=None):
    ix = w.ix
    if deleted is not None:
        ix = O.Index(w.ix.g, w.ix.N, w.ix.P, w.ix.min_key, w.ix.max_key, w.ix.rep, w.ix.ids, deleted)
        store = O.Store(store.dim, store.iv, store.ct, store.key_version, store.keys, deleted)
    ctx.touched(clear=True)
    got = ctx.search_batch(w.queries, k, probes, hard_cap, B)
    codes = O.tokengen_batch(w.queries, w.g)
    touched = np.zeros(w.cfg.N, dtype=np.uint8)
    for q in range(w.queries.shape[0]):
        ref = O.search(ix, store, w.queries[q], codes[q], k, probes, hard_cap, B, touched=touched)
        n = len(ref["top_ids"])
        assert got["n_ret"][q] == n, q
        assert np.array_equal(got["top_ids"][q, :n], ref["top_ids"])
        assert np.array_equal(got["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))
        c = got["counters"][q]
        assert (c[0], c[1], c[2], c[3], bool(c[4])) == (ref["cand_total"], ref["cand_kept"], ref["cand_decrypted"], ref["returned"], ref["retried"])
    assert np.array_equal(ctx.touched(clear=True), np.nonzero(touched)[0])
    return got


def check_refine_verdicts(ctx, w, store, B=96, k=10):
    codes = O.tokengen_batch(w.queries, w.g)
    Q = codes.shape[0]
    cand = np.full((Q, B), -1, dtype=np.int32)
    nc = np.zeros(Q, dtype=np.int32)
    for q in range(Q):
        ids = O.route(w.ix, codes[q], 5, 20000)[0]
        n = min(B, len(ids))
        cand[q, :n], nc[q] = ids[:n], n
    out = ctx.refine_batch(w.queries, cand, nc, k)
    seen = set()
    for q in range(Q):
        ref = O.refine(store, w.queries[q], cand[q, :nc[q]], k)
        assert np.array_equal(out["verdict"][q, :nc[q]], ref["verdict"]), q
        n = len(ref["top_ids"])
        assert out["n_ret"][q] == n and out["n_decrypted"][q] == ref["n_decrypted"]
        assert np.array_equal(out["top_ids"][q, :n], ref["top_ids"])
        assert np.array_equal(out["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))
        seen |= set(ref["verdict"].tolist())
    return seen


def test_mixed_key_versions_and_retire(world_factory):
    w = world_factory(**W3)
    assert set(np.unique(w.key_version)) == {1, 2, 3}
    ctx = w.gpu_context()
    try:
        check_search(ctx, w, w.store)
        assert check_refine_verdicts(ctx, w, w.store) == {N.V_OK}
        # Retire v2 while records are still bound to it: those candidates become "no key" (decryptError in QSI:265-270)
        ctx.keys_retire(2)
        keys = {v: k for v, k in w.store.keys.items() if v != 2}
        st = O.Store(w.store.dim, w.iv, w.ct, w.key_version, keys)
        assert check_refine_verdicts(ctx, w, st) == {N.V_OK, N.V_NO_KEY}
        check_search(ctx, w, st)
        ctx.keys_set(2, w.store.keys[2])          # key comes back: everything verifies again
        assert check_refine_verdicts(ctx, w, w.store) == {N.V_OK}
    finally:
        ctx.close()


def test_tampered_wrong_version_and_non_finite_records(world_factory):
    w = world_factory(**W1)
    rng = np.random.default_rng(17)
    iv, ct, kv = w.iv.copy(), w.ct.copy(), w.key_version.copy()
    ids = rng.permutation(w.cfg.N)
    flip_ct, flip_tag, flip_iv, wrong_ver, nonfin = (ids[i * 150:(i + 1) * 150] for i in range(5))
    ct[flip_ct, rng.integers(0, 8 * 32, size=150)] ^= 0x01
    ct[flip_tag, 8 * 32 + rng.integers(0, 16, size=150)] ^= 0x80
    iv[flip_iv, rng.integers(0, 12, size=150)] ^= 0x10
    keys = {1: w.store.keys[1], 2: O.kdf(w.master, 2)}
    kv[wrong_ver] = 2                           # header says v2, ciphertext was bound to v1: AAD / key mismatch -> tag failure
    bad = w.base[nonfin].copy()
    bad[np.arange(150), rng.integers(0, 32, size=150)] = np.where(np.arange(150) % 2 == 0, np.nan, np.inf)
    ct[nonfin] = O.encrypt_store(bad, 1, keys[1], iv[nonfin], ids=nonfin.astype(np.int32))
    st = O.Store(32, iv, ct, kv, keys)
    ctx = GpuContext(0)
    try:
        g, ix = w.g, w.ix
        ctx.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        for v, k in keys.items():
            ctx.keys_set(v, k)
        ctx.store_upload(32, iv, ct, kv)
        seen = check_refine_verdicts(ctx, w, st)
        assert {N.V_OK, N.V_TAG_FAIL, N.V_NON_FINITE} <= seen
        check_search(ctx, w, st, k=10, B=64)
        check_search(ctx, w, st, k=100, B=200)   # forces the adaptive retry with many failing candidates
    finally:
        ctx.close()


def test_deleted_points_are_skipped_by_route_and_refine(world_factory):
    w = world_factory(**W1)
    rng = np.random.default_rng(4)
    deleted = (rng.random(w.cfg.N) < 0.2).astype(np.uint8)
    ctx = w.gpu_context()
    try:
        ctx.deleted_set(deleted)
        codes = O.tokengen_batch(w.queries, w.g)
        ix = O.Index(w.ix.g, w.ix.N, w.ix.P, w.ix.min_key, w.ix.max_key, w.ix.rep, w.ix.ids, deleted)
        for general in (0, 1):
            ctx.set_option("route_general", general)
            out = ctx.route_batch(codes, 5, 20000, 128)
            for q in range(codes.shape[0]):
                ids, sc, raw, mc = O.route(ix, codes[q], 5, 20000)
                n = min(128, len(ids))
                assert out["unique"][q] == len(ids) and out["raw_seen"][q] == raw
                assert not deleted[out["cand_ids"][q, :n]].any()
                if mc < 9:
                    assert np.array_equal(out["cand_ids"][q, :n], ids[:n])
        ctx.set_option("route_general", 0)
        check_search(ctx, w, w.store, deleted=deleted)
        # a deleted id handed straight to Refine is "not found" (PIS:718)
        cand = np.array([[int(np.nonzero(deleted)[0][0]), int(np.nonzero(deleted == 0)[0][0])]], dtype=np.int32)
        out = ctx.refine_batch(w.queries[:1], cand, np.array([2], dtype=np.int32), 2)
        assert out["verdict"][0].tolist() == [N.V_NOT_FOUND, N.V_OK]
        ctx.deleted_set(None)
        check_search(ctx, w, w.store)
    finally:
        ctx.close()


def test_error_behaviour_mirrors_reference_exceptions(world_factory):
    w = world_factory(**W1)
    ctx = GpuContext(0)
    try:
        g, ix = w.g, w.ix
        codes = O.tokengen_batch(w.queries[:2], w.g)
        with pytest.raises(N.IllegalStateError):                       # PIS:594 "Index not finalized"
            ctx.route_batch(codes, 5, 20000, 64)
        ctx.gfunctions_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega)
        assert np.array_equal(ctx.tokengen_batch(w.queries[:2]), codes)
        with pytest.raises(N.IllegalStateError):                       # GFunctions alone are not a finalized index
            ctx.route_batch(codes, 5, 20000, 64)
        ctx.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        with pytest.raises(N.IllegalStateError):                       # no record store yet
            ctx.search_batch(w.queries[:2], 10, 5, 20000, 64)
        ctx.keys_set(1, w.store.keys[1])
        ctx.store_upload(g.dim, w.iv, w.ct, w.key_version)
        bad = w.queries[:3].copy()
        bad[1, 5] = np.nan
        with pytest.raises(N.IllegalArgumentError):                    # Coding:357-359 "Vector contains NaN/Inf"
            ctx.tokengen_batch(bad)
        with pytest.raises(N.IllegalArgumentError):
            ctx.search_batch(bad, 10, 5, 20000, 64)
        with pytest.raises(N.IllegalArgumentError):                    # dimension mismatch (SuperFailureModeIT)
            ctx.search_batch(np.zeros((2, g.dim + 1)), 10, 5, 20000, 64)
        with pytest.raises(N.IllegalArgumentError):                    # topK must be > 0 (QTF:65)
            ctx.search_batch(w.queries[:2], 0, 5, 20000, 64)
        with pytest.raises(N.IllegalArgumentError):
            ctx.search_batch(w.queries[:2], 10, 5, 20000, 0)
        with pytest.raises(N.IllegalStateError):                       # production build has no plaintext tap
            ctx.debug_decrypt(np.arange(4))
        empty = ctx.search_batch(np.zeros((0, g.dim)), 10, 5, 20000, 64)
        assert empty["top_ids"].shape == (0, 10)
        out = ctx.search_batch(w.queries[:2], 10, 5, 20000, 64)       # still healthy after all the rejected calls
        assert (out["n_ret"] == 10).all()
    finally:
        ctx.close()


def test_debug_tap_plaintext_bit_exact(world_factory):
    """north_star: decrypted plaintexts bit-exact.  Only the -DFSPANN_DEBUG_TAP build can export them."""
    w = world_factory(**W3)
    ctx = w.gpu_context(debug=True)
    try:
        ids = np.concatenate([np.arange(0, 3000, 13), [2999, 0]]).astype(np.int32)
        pt, ver = ctx.debug_decrypt(ids)
        assert (ver == N.V_OK).all()
        assert np.array_equal(pt.view(np.uint64), w.base[ids].view(np.uint64))
        # the debug build runs the same kernels: spot-check search parity through it too
        check_search(ctx, w, w.store)
    finally:
        ctx.close()


def make_system(w, B=64, k=10, debug=False):
    cfg = SystemConfig(PaperConfig(m=w.g.m, lam=w.g.lam, divisions=w.g.D, tables=w.g.T, seed=13),
                       RuntimeConfig(refinementLimit=B, maxGlobalCandidates=20000))
    return ForwardSecureANNSystem(cfg, w.g.dim, w.master, (w.g.alpha, w.g.r, w.g.omega), iv_seed=w.cfg.base_seed + 5, debug=debug)


def test_facade_flow_matches_reference_search(world_factory):
    """FSA.batchInsert -> finalizeForSearch -> createToken -> QueryServiceImpl.search, on the GPU, equals the oracle."""
    w = world_factory(**W1)
    sys_ = make_system(w)
    try:
        with pytest.raises(N.IllegalStateError):                        # FSA:1675-1678
            sys_.createToken(w.queries[0], 10, w.g.dim)
        sys_.batchInsert(w.base, ivs=w.iv)
        assert np.array_equal(sys_.store_ct, w.ct)                      # same key, IVs, AAD => identical ciphertext
        with pytest.raises(N.IllegalStateError):                        # QSI / PIS:594 before finalize
            sys_.queryService.searchBatch([None])
        sys_.finalizeForSearch()
        r = sys_.index.routing
        assert np.array_equal(r.ids, w.ix.ids) and np.array_equal(r.rep, w.ix.rep) and np.array_equal(r.min_key, w.ix.min_key)
        with pytest.raises(N.IllegalArgumentError):                     # FSA:1680-1688 dimension mismatch
            sys_.createToken(w.queries[0][:-1], 10, w.g.dim - 1)
        with pytest.raises(N.IllegalArgumentError):                     # QTF:65
            sys_.tokenFactory.create(w.queries[0], 0)
        assert sys_.queryService.search(None) == []                     # QSI:102
        codes = O.tokengen_batch(w.queries, w.g)
        tokens = [sys_.createToken(w.queries[q], 10, w.g.dim) for q in range(8)]
        for q, t in enumerate(tokens):
            assert np.array_equal(t.bitCodes.reshape(-1, w.g.W), codes[q]) and t.version == 1 and len(t.encryptedQuery) == 8 * w.g.dim + 16
            cands = sys_.index.lookupCandidatesWithScores(t)
            ids, sc, raw, _ = O.route(w.ix, codes[q], 5, 20000)
            assert [c[0] for c in cands] == ids.tolist() and [c[1] for c in cands] == sc.tolist()      # every unique candidate (PIS:690-696)
            assert sys_.index.getLastRawCandidateCount() == raw
            assert sys_.index.getLastTouchedCount() == len(ids) and sys_.index.getLastTouchedIds() == ids.tolist()   # PIS:698-702
            assert sys_.index.lookupCandidateIds(t, limit=64) == ids[:64].tolist()
        results = sys_.queryService.searchBatch(tokens)
        for q, res in enumerate(results):
            ref = O.search(w.ix, w.store, w.queries[q], codes[q], 10, 5, 20000, 64)
            assert [int(r_.id) for r_ in res] == ref["top_ids"].tolist()
            assert [r_.distance for r_ in res] == ref["top_dist"].tolist()
        one = sys_.queryService.search(tokens[3])
        assert [int(r_.id) for r_ in one] == [int(r_.id) for r_ in results[3]]
        assert sys_.queryService.getLastCandDecrypted() == 64 and sys_.queryService.getLastReturned() == 10
    finally:
        sys_.shutdown()


def test_rotate_migrate_retire_keep_results_invariant(world_factory):
    """Config 5 / ARCHITECTURE routing-ciphertext orthogonality: results are identical before Rotate, after Rotate + partial
    Migrate (mixed versions), after a second Rotate + Migrate of the touched set; only listed records change; Retire is
    refused while a version is still bound and then removes the key (KM:287-294)."""
    w = world_factory(**W1)
    sys_ = make_system(w)
    try:
        sys_.batchInsert(w.base, ivs=w.iv)
        sys_.finalizeForSearch()
        tokens = [sys_.createToken(w.queries[q], 10, w.g.dim) for q in range(w.queries.shape[0])]

        def run():
            res = sys_.queryService.searchBatch(tokens)
            return [[(r.id, r.distance) for r in rs] for rs in res], set(sys_.queryService.touchedThisSession)
        base_res, touched1 = run()
        assert len(touched1) > 0
        ct_before = sys_.store_ct.copy()
        # Rotate -> v2, Migrate ids = 0 (mod 3)
        assert sys_.rotateKeyOnly() == 2
        ids = np.arange(0, w.cfg.N, 3, dtype=np.int32)
        done = sys_.reencryptTouched(ids, WL.record_ivs(len(ids), 501), 2)
        assert len(done) == len(ids)
        changed = np.any(sys_.store_ct != ct_before, axis=1)
        assert changed[ids].all() and not changed[np.setdiff1d(np.arange(w.cfg.N), ids)].any()
        res2, touched2 = run()
        assert res2 == base_res and touched2 == touched1
        assert not sys_.retire(1)                                       # still bound to v1
        # old tokens were encrypted under v1: still decryptable (QSI:124-129); new tokens use v2
        assert sys_.createToken(w.queries[0], 10, w.g.dim).version == 2
        # Rotate -> v3, Migrate the touched set of the last batch
        assert sys_.rotateKeyOnly() == 3
        tl = np.array(sorted(touched2), dtype=np.int32)
        sys_.reencryptTouched(tl, WL.record_ivs(len(tl), 502), 3)
        res3, _ = run()
        assert res3 == base_res
        assert set(np.unique(sys_.store_ver)) == {1, 2, 3}
        # migrate everything left on v1, then v1 can be retired and results still hold
        rest = np.nonzero(sys_.store_ver == 1)[0].astype(np.int32)
        sys_.reencryptTouched(rest, WL.record_ivs(len(rest), 503), 3)
        tokens = [sys_.createToken(w.queries[q], 10, w.g.dim) for q in range(w.queries.shape[0])]   # v1 query tokens would die with the key
        assert sys_.retire(1)
        res4, _ = run()
        assert res4 == base_res
        # the oracle agrees on the final mixed-version store
        st = O.Store(w.g.dim, sys_.store_iv, sys_.store_ct, sys_.store_ver, {2: sys_.keys.derive(2), 3: sys_.keys.derive(3)})
        codes = O.tokengen_batch(w.queries, w.g)
        for q in range(0, w.queries.shape[0], 5):
            ref = O.search(w.ix, st, w.queries[q], codes[q], 10, 5, 20000, 64)
            assert [int(i) for i, _ in res4[q]] == ref["top_ids"].tolist()
    finally:
        sys_.shutdown()


def test_large_batch_with_a_non_finite_query_is_rejected_by_the_device_check(world_factory):
    """Batches above 64k values skip the host scan: the device pass that compacts the queries also checks isValid (QSI:407-413);
    the call fails like createToken would (Coding:357-359); the output buffers are unspecified then."""
    w = world_factory(**W1)
    ctx = w.gpu_context()
    try:
        big = np.tile(w.queries, (60, 1))                                # 2400 x 32 = 76800 values
        assert big.size > 65536
        ok = ctx.search_batch(big, 10, 5, 20000, 64)
        assert (ok["n_ret"] == 10).all()
        for poison in (np.nan, np.inf, -np.inf):
            bad = big.copy()
            bad[1777, 13] = poison
            with pytest.raises(N.IllegalArgumentError):
                ctx.search_batch(bad, 10, 5, 20000, 64)
        again = ctx.search_batch(big, 10, 5, 20000, 64)                  # the context is still usable
        assert np.array_equal(again["top_ids"], ok["top_ids"])
    finally:
        ctx.close()


def test_run_queries_max_k_token_fallback_and_per_k_recall(world_factory):
    """FSA.runQueries (FSA:622-748): one token at MAX_K, per-K prefix metrics against ground truth, and the probes-only fallback
    max(2p, 4) for queries whose first search came back empty (here: every candidate of the first pass is deleted)."""
    w = world_factory(**W1)
    sys_ = make_system(w)
    try:
        sys_.batchInsert(w.base, ivs=w.iv)
        sys_.finalizeForSearch()
        kv = (1, 10, 20)
        gt = sys_.gpu.groundtruth(w.base.astype(np.float32), w.queries.astype(np.float32), max(kv))
        out = sys_.runQueries(w.queries, w.g.dim, gt, kv)
        assert out["fallback"] == [] and all(len(r) == max(kv) for r in out["results"])
        codes = O.tokengen_batch(w.queries, w.g)
        for q in range(w.queries.shape[0]):
            ref = O.search(w.ix, w.store, w.queries[q], codes[q], max(kv), 5, 20000, sys_.cfg.runtime.refinementLimit)
            assert [int(r.id) for r in out["results"][q]] == ref["top_ids"].tolist()
            for k in kv:                                              # prefix metrics: recall@K on the first K of the MAX_K result
                got = O.recall_at_k(gt[q, :k], ref["top_ids"], min(k, len(ref["top_ids"])), k)
                assert 0.0 <= got <= 1.0
        for k in kv:
            ref_mean = np.mean([O.recall_at_k(gt[q, :k], np.array([int(r.id) for r in out["results"][q]], dtype=np.int32), min(k, max(kv)), k)
                                for q in range(w.queries.shape[0])])
            assert abs(out["recall"][k] - ref_mean) < 1e-12 and out["returned"][k] == k
        assert sys_.index.effectiveMaxProbes() == 5                    # overrides cleared (FSA:745-746)
        # fallback: delete everything query 0 reaches with 5 probes; with max(2*5, 4) = 10 probes it finds other candidates
        first = sys_.index.lookupCandidatesWithScores(sys_.createToken(w.queries[0], 10, w.g.dim), limit=100000)
        deleted = np.zeros(w.cfg.N, dtype=np.uint8)
        deleted[[i for i, _ in first]] = 1
        sys_.gpu.deleted_set(deleted)
        out2 = sys_.runQueries(w.queries[:3], w.g.dim, None, (10,))
        assert 0 in out2["fallback"]
        ixd = O.Index(w.ix.g, w.ix.N, w.ix.P, w.ix.min_key, w.ix.max_key, w.ix.rep, w.ix.ids, deleted)
        std = O.Store(w.store.dim, w.store.iv, w.store.ct, w.store.key_version, w.store.keys, deleted)
        ref10 = O.search(ixd, std, w.queries[0], codes[0], 10, 10, 20000, sys_.cfg.runtime.refinementLimit)
        assert [int(r.id) for r in out2["results"][0]] == ref10["top_ids"].tolist()
        assert sys_.evalSimple(w.queries[1], 10, w.g.dim) == out2["results"][1]
    finally:
        sys_.gpu.deleted_set(None)
        sys_.shutdown()


def test_byte_distance_path_and_its_fallback_are_bit_exact(world_factory):
    """Integer-valued records (SIFT / .bvecs) take the integer distance path (VABSDIFF4 + DP4A); any record holding a value that is not
    an integer in [0, 255] (fraction, negative, 256, -0.0, denormal) must fall back to the sequential FP64 loop.  Both against the oracle."""
    w = world_factory(**W1)
    base = w.base.copy()
    rng = np.random.default_rng(99)
    odd = rng.choice(w.cfg.N, size=600, replace=False)
    vals = [0.5, -3.0, 256.0, -0.0, 5e-324, 254.99999999999997, 1e300, 255.0, 0.0, 128.0]
    for n_, i in enumerate(odd):
        base[i, rng.integers(0, base.shape[1])] = vals[n_ % len(vals)]
    ct = O.encrypt_store(base, 1, w.store.keys[1], w.iv)
    st = O.Store(w.g.dim, w.iv, ct, w.key_version, dict(w.store.keys))
    ctx = w.gpu_context()
    try:
        ctx.store_upload(w.g.dim, w.iv, ct, w.key_version)
        codes = O.tokengen_batch(w.queries, w.g)
        Q, B = w.queries.shape[0], 96
        cand = np.full((Q, B), -1, dtype=np.int32)
        for q in range(Q):                                            # half routed candidates, half of the doctored records
            ids = O.route(w.ix, codes[q], 5, 20000)[0][:B // 2]
            extra = rng.choice(odd, size=B - len(ids), replace=False)
            cand[q] = np.concatenate([ids, extra]).astype(np.int32)
        nc = np.full(Q, B, dtype=np.int32)
        for queries in (w.queries, w.queries + 0.25):                 # byte-exact queries, then queries that disable the byte path
            out = ctx.refine_batch(queries, cand, nc, 10)
            for q in range(Q):
                ref = O.refine(st, queries[q], cand[q], 10)
                assert np.array_equal(out["verdict"][q], ref["verdict"])
                n = len(ref["top_ids"])
                assert out["n_ret"][q] == n and np.array_equal(out["top_ids"][q, :n], ref["top_ids"])
                assert np.array_equal(out["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))
    finally:
        ctx.close()


def test_small_batch_graph_replay_is_bit_exact_and_follows_state_changes(world_factory):
    """Batches of <= 64 queries replay their first pass as a captured CUDA graph (abi.cu search_core).  The replay must read the
    CURRENT queries, and a deletion / key change between calls must invalidate it: every call is checked against the oracle."""
    w = world_factory(**W3)
    ctx = w.gpu_context()
    try:
        rng = np.random.default_rng(12)
        for Q in (1, 8, 40):
            for rep in range(5):
                w.queries = np.ascontiguousarray(w.base[rng.choice(w.cfg.N, Q, replace=False)] + rng.normal(0, 0.05, (Q, w.cfg.dim)))
                check_search(ctx, w, w.store)
        print("graph captures/replays:", ctx.get_info("graph_captures"), ctx.get_info("graph_replays"))
        assert ctx.get_info("graph_captures") >= 3 and ctx.get_info("graph_replays") >= 6
        # state changes between identical calls
        r0 = ctx.get_info("graph_replays")
        deleted = (rng.random(w.cfg.N) < 0.3).astype(np.uint8)
        ctx.deleted_set(deleted)
        check_search(ctx, w, w.store, deleted=deleted)
        ctx.deleted_set(None)
        ctx.keys_retire(2)
        st = O.Store(w.store.dim, w.iv, w.ct, w.key_version, {v: k for v, k in w.store.keys.items() if v != 2})
        for rep in range(3):
            check_search(ctx, w, st)
        ctx.keys_set(2, w.store.keys[2])
        for rep in range(3):
            check_search(ctx, w, w.store)
        assert ctx.get_info("graph_replays") > r0
        # graphs off: same answers, no replays
        ctx.set_option("graphs", 0)
        r1 = ctx.get_info("graph_replays")
        for rep in range(3):
            check_search(ctx, w, w.store)
        assert ctx.get_info("graph_replays") == r1
        # a batch whose queries force the adaptive retry (far from everything) through the replayed path
        ctx.set_option("graphs", 1)
        w.queries = np.ascontiguousarray(w.queries * 50.0 + 1000.0)
        for rep in range(4):
            check_search(ctx, w, w.store)
    finally:
        ctx.close()
