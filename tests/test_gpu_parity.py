"""Parity of the CUDA path (through the C ABI) with the oracle on identical inputs.  Bit-exact everywhere:
routing codes, ordered candidate lists + scores, raw/unique counters, verdicts, top-k ids AND distances."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

SMALL = dict(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2)
WORLDS = {
    "small": SMALL,
    "sift128": dict(N=5000, dim=128, Q=48, T=4, D=8, m=24, lam=2),
    "glove100": dict(N=4000, dim=100, Q=32, T=4, D=4, m=22, lam=2, shape="glove"),
    "deep96": dict(N=4000, dim=96, Q=32, T=2, D=4, m=24, lam=2, shape="deep"),
    "odd-dim": dict(N=2000, dim=33, Q=24, T=2, D=3, m=10, lam=3, shape="glove"),
    "wide-code": dict(N=2500, dim=24, Q=24, T=2, D=2, m=24, lam=3, shape="glove"),   # 72 code bits: W=2
    "m30": dict(N=2500, dim=20, Q=16, T=2, D=2, m=30, lam=2, shape="glove"),          # m > 24: TokenGen takes the exact FP64 kernel
}


@pytest.fixture(scope="module", params=list(WORLDS))
def wg(request, world_factory):
    w = world_factory(**WORLDS[request.param])
    ctx = w.gpu_context()
    yield w, ctx
    ctx.close()


def test_tokengen_codes_bit_exact(wg):
    w, ctx = wg
    ref = O.tokengen_batch(w.queries, w.g)
    got = ctx.tokengen_batch(w.queries)
    assert got.dtype == np.uint64 and got.shape == ref.shape
    assert np.array_equal(got, ref)
    # Setup-side coding of base vectors goes through the same kernel
    assert np.array_equal(ctx.tokengen_batch(w.base[:777]), w.codes[:777])


@pytest.mark.parametrize("general", [0, 2, 1], ids=["fast-path", "fast-path-one-cta", "general-path"])
@pytest.mark.parametrize("probes,hard_cap,B", [(5, 20000, 64), (5, 20000, 100000), (3, 700, 256), (10, 2000, 1500), (1, 50, 40), (5, 1 << 20, 300),
                                               (5, 24000, 1024), (2, 20000, 1)])
def test_route_candidates_ordered_bit_exact(wg, probes, hard_cap, B, general):
    """Both Route kernels: the shared-memory fast path (when the HARD_CAP cannot bind and the query fits in smem) and the
    general path (sequential groups, exact cap semantics) must give the reference's ordered candidate list."""
    w, ctx = wg
    codes = O.tokengen_batch(w.queries, w.g)
    hard_cap = max(hard_cap, min(B, 5000)) if hard_cap < B and B < 100000 else hard_cap
    B = min(B, 8192)
    ctx.set_option("route_general", int(general == 1))
    ctx.set_option("route_v1", int(general == 2))
    ctx.set_option("route_small_v1", 0)         # (the fixture's setting) batches smaller than the SM count stay on the two-CTA kernel under test
    try:
        out = ctx.route_batch(codes, probes, hard_cap, B)
    finally:
        ctx.set_option("route_general", 0)
        ctx.set_option("route_v1", 0)
    if general == 1:
        assert ctx.get_info("last_route_path") == 2
    elif ctx.get_info("last_route_path") == 1:
        # fast path: the two-CTA kernel serves B <= 1024 unless forced off; queries it handed to the one-CTA kernel are counted
        # (the one-CTA kernel needs a HARD_CAP that cannot bind and B small enough for its shared memory; forced, it yields to the two-CTA
        #  kernel where it is not eligible)
        assert ctx.get_info("last_route_v2") == 1 if general == 0 else ctx.get_info("last_route_v2") in (0, 1)
        assert 0 <= ctx.get_info("route_overflowed") <= codes.shape[0]
        if general == 0 and w.g.T * w.g.D * probes * 64 <= 8192:       # short position lists fit the two-CTA kernel's worklist entirely
            assert ctx.get_info("route_overflowed") == 0
    general = int(general == 1)
    worst_chain = 0
    for q in range(codes.shape[0]):
        ids, sc, raw, mc = O.route(w.ix, codes[q], probes, hard_cap)
        worst_chain = max(worst_chain, mc)
        n = min(B, len(ids))
        assert out["n_cand"][q] == n
        assert out["unique"][q] == len(ids)
        assert out["raw_seen"][q] == raw
        assert np.array_equal(out["cand_scores"][q, :n], sc[:n])
        if mc >= 9:
            # a bin of the Java HashMap reached 9 entries: java.util.HashMap would treeify it and the iteration order inside
            # that bin is no longer insertion order.  Neither the oracle nor the kernel models tree bins (limits);
            # counts, scores and (when nothing is cut) the candidate set are still order-independent.
            if n == len(ids):
                assert set(out["cand_ids"][q, :n].tolist()) == set(ids.tolist())
            continue
        assert np.array_equal(out["cand_ids"][q, :n], ids[:n]), f"query {q}"
    flag = ctx.get_info("route_treeified")
    if general and worst_chain >= 9:   # the general kernel reports a bin of 9 (it checks at the map's initial capacity: it may flag more than the oracle's final-capacity audit, never less)
        assert flag == 1
    assert flag in (0, 1) and (flag == 0 or ctx.get_info("last_route_path") == 2)


def test_refine_verdicts_topk_bit_exact(wg):
    w, ctx = wg
    codes = O.tokengen_batch(w.queries, w.g)
    Q, B, k = codes.shape[0], 128, 10
    cand = np.full((Q, B), -1, dtype=np.int32)
    ncand = np.zeros(Q, dtype=np.int32)
    rng = np.random.default_rng(3)
    for q in range(Q):
        ids, _, _, _ = O.route(w.ix, codes[q], 5, 20000)
        n = min(B, len(ids)) if q % 5 else min(B, len(ids)) // 2     # ragged lists
        if q == 7:
            n = 0                                                      # empty list
        cand[q, :n] = ids[:n]
        if q == 3 and n > 4:
            cand[q, 1] = w.cfg.N + 5                                   # unknown id -> notFound
            cand[q, 2] = -7
        ncand[q] = n
    out = ctx.refine_batch(w.queries, cand, ncand, k)
    for q in range(Q):
        ref = O.refine(w.store, w.queries[q], cand[q, :ncand[q]], k)
        n = len(ref["top_ids"])
        assert out["n_ret"][q] == n
        assert out["n_decrypted"][q] == ref["n_decrypted"]
        assert np.array_equal(out["verdict"][q, :ncand[q]], ref["verdict"])
        assert np.array_equal(out["top_ids"][q, :n], ref["top_ids"])
        assert np.array_equal(out["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64)), "distances must be bit-exact FP64"


@pytest.mark.parametrize("B", [2500, 600], ids=["radix-select", "warp-kernel"])
@pytest.mark.parametrize("k", [1, 10, 16, 17, 64, 300])
def test_refine_long_lists_with_exact_distance_ties(wg, k, B):
    """Candidate lists beyond 1024 entries take the radix-select top-k.  Lists that name the same records several times give many
    EXACT distance ties, which the reference's stable sort (QSI:298) resolves by candidate order; ragged / empty lists included."""
    w, ctx = wg
    Q = min(12, w.queries.shape[0])
    rng = np.random.default_rng(17 + k)
    cand = np.full((Q, B), -1, dtype=np.int32)
    ncand = np.zeros(Q, dtype=np.int32)
    for q in range(Q):
        n = [B, 1100, 0, 1025, 2499][q % 5] if B > 1024 else [B, 33, 0, 1, B - 1][q % 5]
        distinct = rng.choice(w.cfg.N, size=min(w.cfg.N, max(1, n // (1 + q % 4))), replace=False)
        lst = rng.choice(distinct, size=n, replace=True) if n else np.zeros(0, dtype=np.int64)
        cand[q, :n] = lst
        if n > 10:
            cand[q, 5] = w.cfg.N + 3                                    # notFound among them
        ncand[q] = n
    out = ctx.refine_batch(w.queries[:Q], cand, ncand, k)
    for q in range(Q):
        ref = O.refine(w.store, w.queries[q], cand[q, :ncand[q]], k)
        n = len(ref["top_ids"])
        assert out["n_ret"][q] == n and out["n_decrypted"][q] == ref["n_decrypted"]
        assert np.array_equal(out["top_ids"][q, :n], ref["top_ids"]), q
        assert np.array_equal(out["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))


@pytest.mark.parametrize("k,B", [(10, 64), (100, 256), (1, 16), (100, 3000), (7, 1500)])     # B > 1024: radix-select top-k
def test_search_batch_matches_reference_search(wg, k, B):
    """QSI.search incl. the adaptive retry (k=100,B=256 forces decrypted < 10*K -> second pass with 10 probes)."""
    w, ctx = wg
    hard_cap = 20000
    ctx.touched(clear=True)
    got = ctx.search_batch(w.queries, k, 5, hard_cap, B)
    touched_ref = np.zeros(w.cfg.N, dtype=np.uint8)
    codes = O.tokengen_batch(w.queries, w.g)
    for q in range(w.queries.shape[0]):
        ref = O.search(w.ix, w.store, w.queries[q], codes[q], k, 5, hard_cap, B, touched=touched_ref)
        n = len(ref["top_ids"])
        assert got["n_ret"][q] == n
        assert np.array_equal(got["top_ids"][q, :n], ref["top_ids"])
        assert np.array_equal(got["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))
        c = got["counters"][q]
        assert (c[0], c[1], c[2], c[3], bool(c[4])) == (ref["cand_total"], ref["cand_kept"], ref["cand_decrypted"], ref["returned"], ref["retried"])
    assert np.array_equal(ctx.touched(clear=True), np.nonzero(touched_ref)[0])


def test_distance_paths_fp32_exact_and_general_fp64(wg):
    """The distance phase reads a compact FP32 copy of the queries when every value is FP32-representable (the reference's
    loaders widen float32 input), otherwise the FP64 rows: both must give the oracle's FP64 bits."""
    w, ctx = wg
    rng = np.random.default_rng(8)
    assert np.array_equal(w.queries.astype(np.float32).astype(np.float64), w.queries)
    q64 = w.queries + rng.normal(0, 1e-9, size=w.queries.shape)          # not representable in FP32
    assert not np.array_equal(q64.astype(np.float32).astype(np.float64), q64)
    for queries in (w.queries, q64):
        got = ctx.search_batch(queries, 10, 5, 20000, 128)
        codes = O.tokengen_batch(queries, w.g)
        for q in range(queries.shape[0]):
            ref = O.search(w.ix, w.store, queries[q], codes[q], 10, 5, 20000, 128)
            n = len(ref["top_ids"])
            assert got["n_ret"][q] == n and np.array_equal(got["top_ids"][q, :n], ref["top_ids"])
            assert np.array_equal(got["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))


@pytest.mark.parametrize("wl_extra", [-1, 0], ids=["worklist", "worklist-overflow-fallback"])
def test_route_fast_path_with_every_id_duplicated(world_factory, wl_extra):
    """Correlated tables: every table reuses table 0's projections (own offsets), so most ids are visited several times per
    query with different scores and most positions take the fast path's exact (hash + chain) route.  wl_extra=0 clamps the worklist so it overflows and the kernel must fall
    back to scanning every position.  Order, scores and raw/unique counters must still be the reference's."""
    import copy
    base_w = world_factory(**SMALL)
    g = copy.deepcopy(base_w.g)
    TD, D = g.T * g.D, g.D
    for td in range(D, TD):
        g.alpha[td], g.omega[td] = g.alpha[td % D], g.omega[td % D]     # same projections, different offsets r
    codes_base = O.tokengen_batch(base_w.base, g)
    ix = O.index_build(codes_base, g, O.staged_order(base_w.cfg.N))
    from fspann_query_system_b200.gpu import GpuContext
    ctx = GpuContext(0)
    try:
        ctx.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        ctx.set_option("route_wl_extra", wl_extra)
        codes = O.tokengen_batch(base_w.queries, g)
        for probes, B in [(5, 64), (8, 700)]:
            out = ctx.route_batch(codes, probes, 1 << 20, B)
            assert ctx.get_info("last_route_path") == 1
            if wl_extra == 0 and ctx.get_info("last_route_v2"):          # two-CTA kernel with a 16-entry worklist: queries overflow into the one-CTA kernel
                assert ctx.get_info("route_overflowed") > 0
            tot_raw = tot_unique = 0
            for q in range(codes.shape[0]):
                ids, sc, raw, mc = O.route(ix, codes[q], probes, 1 << 20)
                n = min(B, len(ids))
                tot_raw += raw; tot_unique += len(ids)
                assert (out["n_cand"][q], out["unique"][q], out["raw_seen"][q]) == (n, len(ids), raw)
                assert np.array_equal(out["cand_scores"][q, :n], sc[:n])
                if mc < 9:
                    assert np.array_equal(out["cand_ids"][q, :n], ids[:n]), f"query {q}"
            assert tot_raw > tot_unique                                  # strict improvements (so: duplicates) really occur
            assert tot_unique < 0.8 * codes.shape[0] * g.T * g.D * probes * 64
    finally:
        ctx.close()


def test_routing_build_on_device_matches_reference_build(wg):
    """finalizeForSearch on the device (coding loop + GreedyPartitioner.build per division): partition keys,
    representative codes and id order are the reference's bit for bit, and the installed state routes like the oracle."""
    from fspann_query_system_b200 import _native as N
    from fspann_query_system_b200.gpu import GpuContext
    w, _ = wg
    g, ix, n = w.g, w.ix, w.cfg.N
    ctx = GpuContext(0)
    try:
        with pytest.raises(N.IllegalStateError):                         # no GFunctions yet (PIS:812-819)
            ctx.dim, ctx.T, ctx.D, ctx.W = g.dim, g.T, g.D, g.W
            ctx.routing_build(w.base, O.staged_order(n))
        ctx.gfunctions_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega)
        with pytest.raises(N.IllegalStateError):                         # PIS:803-808
            ctx.routing_build(w.base[:999], O.staged_order(1000)[:999] % 999)
        bad = O.staged_order(n).copy(); bad[5] = bad[6]
        with pytest.raises(N.IllegalArgumentError):
            ctx.routing_build(w.base, bad)
        mn, mx, rep, ids = ctx.routing_build(w.base, O.staged_order(n))
        assert ctx.get_info("build_treeified") == (1 if ix.max_chain >= 9 else 0)
        assert np.array_equal(ids, ix.ids), "partition order of ids"
        assert np.array_equal(mn, ix.min_key) and np.array_equal(mx, ix.max_key)
        assert np.array_equal(rep, ix.rep)
        codes = O.tokengen_batch(w.queries, g)
        out = ctx.route_batch(codes, 5, 20000, 128)
        for q in range(0, codes.shape[0], 3):
            rids, sc, raw, mc = O.route(ix, codes[q], 5, 20000)
            k = min(128, len(rids))
            assert out["n_cand"][q] == k and out["raw_seen"][q] == raw and np.array_equal(out["cand_scores"][q, :k], sc[:k])
            if mc < 9:
                assert np.array_equal(out["cand_ids"][q, :k], rids[:k])
        # a different insertion order gives the reference's (different) HashMap iteration order too
        perm = np.random.default_rng(4).permutation(n).astype(np.int32)
        ix2 = O.index_build(w.codes, g, perm)
        mn2, mx2, rep2, ids2 = ctx.routing_build(w.base, perm)
        assert np.array_equal(ids2, ix2.ids) and np.array_equal(rep2, ix2.rep) and np.array_equal(mn2, ix2.min_key) and np.array_equal(mx2, ix2.max_key)
    finally:
        ctx.close()


def test_tokengen_prefilter_and_exact_kernel_agree_incl_boundary_vectors(wg):
    """TokenGen runs a reduced-precision pre-filter (tensor cores: BF16-split tcgen05.mma with the accumulator in TMEM; or the FP32 FMA
    pipe) with an error bound and re-checks the undecided projections with the exact sequential FP64 arithmetic.  Codes must equal the exact kernel's and the oracle's: on ordinary data, on vectors constructed to sit ON
    quantisation boundaries ((alpha.v + r)/omega within 1e-12 of an integer -> must be re-checked), on huge / tiny magnitudes, and
    when the re-check list overflows (whole batch recomputed by the exact kernel)."""
    w, ctx = wg
    g = w.g
    rng = np.random.default_rng(12)
    base = w.base[:1500]
    # boundary vectors: scale a data vector so that projection (td, j) lands exactly on k*omega - r
    tricky = []
    for s in range(200):
        v = base[s].copy()
        td, j = int(rng.integers(0, g.T * g.D)), int(rng.integers(0, g.m))
        y = float(v @ g.alpha[td, j])
        if abs(y) < 1e-6:
            continue
        k = np.floor((y + g.r[td, j]) / g.omega[td, j]) + rng.integers(0, 2)
        target = k * g.omega[td, j] - g.r[td, j]
        tricky.append(v * (target / y) * (1.0 + rng.choice([0.0, 1e-15, -1e-15, 3e-13, -3e-13])))
    extremes = np.stack([base[0] * 1e30, base[1] * 1e-30, base[2] * 1e300, np.zeros(g.dim), -base[3]])
    vecs = np.concatenate([base, np.asarray(tricky), extremes])
    ref = O.tokengen_batch(vecs, g)
    total = vecs.shape[0] * g.T * g.D * g.m
    paths = set()
    try:
        for mode in (0, 2):                                              # 0: tensor-core pre-filter (tcgen05) where the shape allows, 2: FP32 pre-filter
            ctx.set_option("tokengen_mode", mode)
            got = ctx.tokengen_batch(vecs)
            n_re, path = ctx.get_info("tokengen_rechecked"), ctx.get_info("last_tokengen_path")
            paths.add(path)
            assert np.array_equal(got, ref), (mode, path)
            if path != 1:                                                # (path 1 = the exact kernel alone: nothing to re-check)
                assert len(tricky) <= n_re < 0.2 * total, (n_re, total)  # the boundary cases were re-checked; the bulk was not
            assert ctx.get_info("tokengen_overflow") == 0
            ctx.set_option("tokengen_list_cap", 8)                       # overflow -> the exact kernel recomputes the batch
            assert np.array_equal(ctx.tokengen_batch(vecs), ref)
            assert ctx.get_info("tokengen_overflow") == (1 if path != 1 else 0)
            ctx.set_option("tokengen_list_cap", 0)
        ctx.set_option("tokengen_exact", 1)
        assert np.array_equal(ctx.tokengen_batch(vecs), ref) and ctx.get_info("last_tokengen_path") == 1
    finally:
        ctx.set_option("tokengen_exact", 0)
        ctx.set_option("tokengen_mode", 0)
        ctx.set_option("tokengen_list_cap", 0)
    if g.dim <= 128 and g.W == 1:
        assert 3 in paths                                                # the tensor-core path did run for this shape
    assert np.array_equal(ctx.tokengen_batch(vecs[:7]), ref[:7]) and ctx.get_info("tokengen_overflow") == 0


def test_small_batches_take_the_one_cta_route_kernel_with_identical_results(wg):
    """Batches of at most one query per SM route on the one-CTA kernel by default (latency); same candidates as the two-CTA kernel."""
    w, ctx = wg
    codes = O.tokengen_batch(w.queries, w.g)
    a = ctx.route_batch(codes, 5, 1 << 20, 64)
    va = ctx.get_info("last_route_v2")
    ctx.set_option("route_small_v1", 1)
    try:
        b = ctx.route_batch(codes, 5, 1 << 20, 64)
        vb = ctx.get_info("last_route_v2")
    finally:
        ctx.set_option("route_small_v1", 0)
    if ctx.get_info("last_route_path") == 1:
        assert va == 1 and vb == 0
    for key in ("cand_ids", "cand_scores", "n_cand", "unique", "raw_seen"):
        assert np.array_equal(a[key], b[key]), key
