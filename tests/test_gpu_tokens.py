"""Token-faithful search: QueryServiceImpl.search routes on token.getBitCodes() (PIS:600), it never recomputes the
codes from the decrypted query; QueryTokenFactory.derive (QTF:182-200); a non-finite query returns empty (QSI:137); several contexts
(one per GPU) inside one process."""
import numpy as np
import pytest
import torch

from fspann_query_system_b200 import _native as N, workloads as WL
from fspann_query_system_b200.api import ForwardSecureANNSystem, PaperConfig, RuntimeConfig, SystemConfig
from fspann_query_system_b200.gpu import GpuContext
from oracle import oracle as O

pytestmark = pytest.mark.gpu

W1 = dict(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2)
WBIG = dict(N=20000, dim=128, Q=64, T=4, D=8, m=24, lam=2)


def _same(got, q, ref):
    n = len(ref["top_ids"])
    assert got["n_ret"][q] == n, q
    assert np.array_equal(got["top_ids"][q, :n], ref["top_ids"])
    assert np.array_equal(got["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))
    c = got["counters"][q]
    assert (c[0], c[1], c[2], c[3], bool(c[4])) == (ref["cand_total"], ref["cand_kept"], ref["cand_decrypted"], ref["returned"], ref["retried"])


@pytest.mark.parametrize("wk,k,B", [(W1, 10, 64), (WBIG, 10, 256), (WBIG, 100, 256)])
def test_search_tokens_routes_on_the_supplied_codes(world_factory, wk, k, B):
    """Tokens whose codes are NOT C(q) (here: the codes of the next query): the result must be the oracle's orc_search on THOSE codes
    -- candidates from the foreign codes, distances to the token's own query.  k=100, B=256 forces the 10-probe retry pass, which must
    re-route the same foreign codes."""
    w = world_factory(**wk)
    ctx = w.gpu_context()
    try:
        codes = O.tokengen_batch(w.queries, w.g)
        foreign = np.roll(codes, 1, axis=0)
        assert not np.array_equal(foreign, codes)
        ctx.touched(clear=True)
        got = ctx.search_tokens(foreign, w.queries, k, 5, 20000, B)
        touched = np.zeros(w.cfg.N, dtype=np.uint8)
        retried = 0
        for q in range(w.queries.shape[0]):
            ref = O.search(w.ix, w.store, w.queries[q], foreign[q], k, 5, 20000, B, touched=touched)
            _same(got, q, ref)
            retried += ref["retried"]
        assert np.array_equal(ctx.touched(clear=True), np.nonzero(touched)[0])
        if k == 100:
            assert retried == w.queries.shape[0]
        # with the genuine codes it equals fspann_search_batch (createToken + search), which runs TokenGen itself
        a, b = ctx.search_tokens(codes, w.queries, k, 5, 20000, B), ctx.search_batch(w.queries, k, 5, 20000, B)
        for key in ("top_ids", "n_ret", "counters"):
            assert np.array_equal(a[key], b[key])
        assert np.array_equal(a["top_dist"].view(np.uint64), b["top_dist"].view(np.uint64))
    finally:
        ctx.close()


def test_search_tokens_dev_matches_host_variant(world_factory):
    w = world_factory(**WBIG)
    ctx = w.gpu_context()
    try:
        k, B, Q = 10, 256, w.queries.shape[0]
        codes = np.roll(O.tokengen_batch(w.queries, w.g), 3, axis=0)
        ref = ctx.search_tokens(codes, w.queries, k, 5, 20000, B)
        dq, dc = torch.from_numpy(w.queries).cuda(), torch.from_numpy(codes.view(np.int64)).cuda()
        ids = torch.empty((Q, k), dtype=torch.int32, device="cuda"); dd = torch.empty((Q, k), dtype=torch.float64, device="cuda")
        nr = torch.empty((Q,), dtype=torch.int32, device="cuda"); cn = torch.empty((Q, 6), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for allow_retry in (1, 0):
            ctx.search_tokens_dev(Q, dc.data_ptr(), dq.data_ptr(), k, 5, 20000, B, 0, allow_retry, ids.data_ptr(), dd.data_ptr(), nr.data_ptr(), cn.data_ptr())
            ctx.sync()
            assert np.array_equal(ids.cpu().numpy(), ref["top_ids"]) and np.array_equal(nr.cpu().numpy(), ref["n_ret"])
            assert np.array_equal(dd.cpu().numpy().view(np.uint64), ref["top_dist"].view(np.uint64))
            assert np.array_equal(cn.cpu().numpy(), ref["counters"])
    finally:
        ctx.close()


def test_non_finite_query_in_a_token_returns_empty_and_touches_nothing(world_factory):
    """QSI:137: a decrypted query with NaN/Inf returns an empty list (the token exists already, nothing throws); the other queries of
    the batch are unaffected and the poisoned one marks no record as touched."""
    w = world_factory(**W1)
    ctx = w.gpu_context()
    try:
        codes = O.tokengen_batch(w.queries, w.g)
        good = ctx.search_tokens(codes, w.queries, 10, 5, 20000, 64)
        for Qn, poison in ((40, np.nan), (2400, np.inf), (2400, -np.inf)):                 # small and > 64k-value batches
            reps = Qn // 40
            qs, cs = np.tile(w.queries, (reps, 1)), np.tile(codes, (reps, 1, 1))
            bad = qs.copy()
            bad[17, 5] = poison
            ctx.touched(clear=True)
            only = ctx.search_tokens(cs[17:18], bad[17:18], 10, 5, 20000, 64)
            assert only["n_ret"][0] == 0 and (only["top_ids"] == -1).all() and (only["counters"] == 0).all()
            assert ctx.touched(clear=True).size == 0
            out = ctx.search_tokens(cs, bad, 10, 5, 20000, 64)
            assert out["n_ret"][17] == 0 and (out["top_ids"][17] == -1).all()
            keep = np.arange(Qn) != 17
            assert np.array_equal(out["top_ids"][keep], np.tile(good["top_ids"], (reps, 1))[keep])
            assert (out["n_ret"][keep] == 10).all()
        with pytest.raises(N.IllegalStateError):                                        # PIS:604-606 "QueryToken missing BitSet codes"
            ctx.search_tokens(None, w.queries, 10, 5, 20000, 64)
        # the refine entry applies the same rule
        bad = w.queries.copy(); bad[3, 0] = np.nan
        r = ctx.route_batch(codes, 5, 20000, 64)
        f = ctx.refine_batch(bad, r["cand_ids"], r["n_cand"], 10)
        assert f["n_ret"][3] == 0 and f["n_decrypted"][3] == 0 and (f["n_ret"][np.arange(40) != 3] == 10).all()
    finally:
        ctx.close()


def _system(w, B=64, iv_seed=7):
    cfg = SystemConfig(PaperConfig(m=w.g.m, lam=w.g.lam, divisions=w.g.D, tables=w.g.T, seed=13), RuntimeConfig(refinementLimit=B, maxGlobalCandidates=20000))
    return ForwardSecureANNSystem(cfg, w.g.dim, w.master, (w.g.alpha, w.g.r, w.g.omega), iv_seed=iv_seed)


def test_facade_search_honours_token_codes_and_derive(world_factory):
    """QueryServiceImpl.search(token) uses token.getBitCodes(): swapping the codes of two tokens swaps their candidate sets.  derive()
    keeps codes / IV / ciphertext and only changes topK (QTF:182-200)."""
    w = world_factory(**W1)
    sys_ = _system(w)
    try:
        sys_.batchInsert(w.base, ivs=w.iv)
        sys_.finalizeForSearch()
        codes = O.tokengen_batch(w.queries, w.g)
        t0, t1 = sys_.createToken(w.queries[0], 10, w.g.dim), sys_.createToken(w.queries[1], 10, w.g.dim)
        t0.bitCodes, t1.bitCodes = t1.bitCodes, t0.bitCodes
        res = sys_.queryService.searchBatch([t0, t1])
        for q, other in ((0, 1), (1, 0)):
            ref = O.search(w.ix, w.store, w.queries[q], codes[other], 10, 5, 20000, 64)
            assert [int(r.id) for r in res[q]] == ref["top_ids"].tolist() and [r.distance for r in res[q]] == ref["top_dist"].tolist()
        t = sys_.createToken(w.queries[2], 10, w.g.dim)
        d = sys_.tokenFactory.derive(t, 3)
        assert d.topK == 3 and np.array_equal(d.bitCodes, t.bitCodes) and d.iv == t.iv and d.encryptedQuery == t.encryptedQuery and d.version == t.version
        assert [r.id for r in sys_.queryService.search(d)] == [r.id for r in sys_.queryService.search(t)][:3]
        with pytest.raises(N.IllegalArgumentError):
            sys_.tokenFactory.derive(t, 0)
        with pytest.raises(N.IllegalArgumentError):
            sys_.tokenFactory.derive(None, 5)
        # clearProbeOverride runs in `finally` (QSI:342-346): a search that throws still clears it
        sys_.index.setProbeOverride(7)
        broken = sys_.createToken(w.queries[0], 10, w.g.dim)
        broken.bitCodes = broken.bitCodes[:1]
        with pytest.raises(N.IllegalStateError):
            sys_.queryService.search(broken)
        t_ok = sys_.createToken(w.queries[0], 10, w.g.dim)
        sys_.index.setProbeOverride(7)
        orig = sys_.gpu.search_tokens
        sys_.gpu.search_tokens = lambda *a, **k: (_ for _ in ()).throw(N.CudaError("boom"))
        with pytest.raises(N.CudaError):
            sys_.queryService.search(t_ok)
        sys_.gpu.search_tokens = orig
        assert sys_.index.effectiveMaxProbes() == 5
    finally:
        sys_.shutdown()


def test_facade_default_ivs_are_random_and_second_batch_insert_appends(world_factory):
    """Record and query IVs default to the OS CSPRNG (AGC:66-67, QTF:152-154): two systems with the same master key never share an
    IV.  A second batchInsert continues the ordinals (FSA:501,515) instead of replacing the index."""
    w = world_factory(**W1)
    a, b = _system(w, iv_seed=None), _system(w, iv_seed=None)
    try:
        half = w.cfg.N // 2
        a.batchInsert(w.base[:half]); a.batchInsert(w.base[half:])
        b.batchInsert(w.base)
        assert a.store_iv.shape == (w.cfg.N, 12) and a.store_ver.shape == (w.cfg.N,)
        assert not np.array_equal(a.store_iv, b.store_iv)
        assert len({bytes(r) for r in a.store_iv}) == w.cfg.N                       # no IV repeats under one key
        a.finalizeForSearch(); b.finalizeForSearch()
        assert np.array_equal(a.index.routing.ids, w.ix.ids) and np.array_equal(b.index.routing.ids, w.ix.ids)
        ta, tb = a.createToken(w.queries[0], 10, w.g.dim), b.createToken(w.queries[0], 10, w.g.dim)
        assert ta.iv != tb.iv and ta.encryptedQuery != tb.encryptedQuery
        codes = O.tokengen_batch(w.queries[:1], w.g)
        ra, rb = a.queryService.search(ta), b.queryService.search(tb)
        assert [r.id for r in ra] == [r.id for r in rb]                             # routing / results do not depend on the IVs
        st = O.Store(w.g.dim, a.store_iv, a.store_ct, a.store_ver, {1: O.kdf(w.master, 1)})
        ref = O.search(w.ix, st, w.queries[0], codes[0], 10, 5, 20000, 64)
        assert [int(r.id) for r in ra] == ref["top_ids"].tolist()
    finally:
        a.shutdown(); b.shutdown()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_contexts_on_two_devices_in_one_process(world_factory):
    """One process, one context per GPU (the `one JVM thread per GPU` mode of INTEGRATION.md): the > 48 KB shared-memory opt-in of the
    kernels is a per-device attribute and must be configured for every device a context is created on."""
    import threading
    w = world_factory(**WBIG)
    g, ix = w.g, w.ix
    ctxs = []
    try:
        for dev in (0, 1):
            c = GpuContext(dev)
            c.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
            c.keys_set(1, w.store.keys[1])
            c.store_upload(g.dim, w.store.iv, w.store.ct, w.store.key_version)
            ctxs.append(c)
        outs = [None, None]

        def run(i):
            for _ in range(3):
                outs[i] = ctxs[i].search_batch(w.queries, 10, 5, 20000, 256)
        ts = [threading.Thread(target=run, args=(i,)) for i in range(2)]
        [t.start() for t in ts]; [t.join() for t in ts]
        codes = O.tokengen_batch(w.queries, w.g)
        for q in range(0, w.queries.shape[0], 7):
            ref = O.search(w.ix, w.store, w.queries[q], codes[q], 10, 5, 20000, 256)
            for o in outs:
                _same(o, q, ref)
        assert np.array_equal(outs[0]["top_ids"], outs[1]["top_ids"])
        assert ctxs[1].get_info("sm_count") > 0
    finally:
        for c in ctxs:
            c.close()
