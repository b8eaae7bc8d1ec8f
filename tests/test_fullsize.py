"""BASELINE configs[1] / configs[4] at FULL size (SIFT1M-shape 1M x 128, T=D=8, B=1024) through size-independent properties, plus
an oracle spot check: device Setup (coding + partition build + bulk encryption), search, Rotate + partial Migrate on the device
(config 5: results must not move), ground truth / recall.  One world, ~20 s on a B200 box."""
import numpy as np
import pytest

from fspann_query_system_b200 import hostsetup as HS, workloads as WL
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    from fspann_query_system_b200.gpu import GpuContext
    cfg = WL.CONFIGS["C2"]
    base = WL.base_vectors(cfg)
    alpha, r, omega = HS.build_gfunctions(base[:HS.MIN_SAMPLE_SIZE], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    gpu = GpuContext(0)
    gpu.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
    mn, mx, rep, ids = gpu.routing_build(base, HS.staged_order(cfg.N))
    km = HS.KeyManager(WL.MASTER_KEY)
    iv = WL.record_ivs(cfg.N, cfg.base_seed + 5)
    gpu.keys_set(1, km.derive(1))
    ct = gpu.encrypt_batch(np.arange(cfg.N, dtype=np.int32), base, iv, 1)
    gpu.store_upload(cfg.dim, iv, ct, np.ones(cfg.N, dtype=np.int32))
    queries = WL.query_vectors(cfg, 2000)
    yield dict(cfg=cfg, gpu=gpu, base=base, alpha=alpha, r=r, omega=omega, mn=mn, mx=mx, rep=rep, ids=ids, iv=iv, ct=ct, km=km, queries=queries)
    gpu.close()


def test_device_built_index_is_a_valid_greedy_partitioning(c2):
    cfg, mn, mx, ids = c2["cfg"], c2["mn"], c2["mx"], c2["ids"]
    assert c2["gpu"].get_info("build_treeified") == 0                     # decimal ids up to 1M never fill a HashMap bin with 9 entries
    assert ids.shape == (cfg.T * cfg.D, cfg.N) and mn.shape == (cfg.T * cfg.D, (cfg.N + 63) // 64)
    assert (mn <= mx).all() and (mn[:, 1:] >= mx[:, :-1]).all() and (mn >= 0).all()      # disjoint sorted key ranges (GP:51-72)
    for td in (0, 17, cfg.T * cfg.D - 1):
        assert np.array_equal(np.sort(ids[td]), np.arange(cfg.N, dtype=np.int32))          # every id exactly once per division
    # spot check against the host restatement of GP.build for one division
    codes = c2["gpu"].tokengen_batch(c2["base"][:4096])
    keys = HS.compute_keys(codes[:, 5, 0])
    pos_of = np.empty(cfg.N, dtype=np.int64); pos_of[ids[5]] = np.arange(cfg.N)
    p = pos_of[:4096] // 64
    assert ((mn[5][p] <= keys) & (keys <= mx[5][p])).all()                                  # each id sits in the partition covering its key


def test_tokengen_prefilter_equals_exact_kernel_on_the_whole_base_set(c2):
    """1.5 G projections, twice: the tensor-core pre-filter (tcgen05.mma, BF16-split operands, FP32 accumulation in TMEM) and the FP32
    pre-filter, each followed by the exact re-check, must give the exact kernel's codes for every base vector.  This is also the
    empirical check of the tensor-core error bound (tokengen_tc.cu): millions of these projections sit within 1e-3 of a boundary."""
    gpu, base = c2["gpu"], c2["base"]
    tot_re = {0: 0, 2: 0}
    for s in range(0, base.shape[0], 250_000):
        gpu.set_option("tokengen_exact", 1)
        try:
            exact = gpu.tokengen_batch(base[s:s + 250_000])
        finally:
            gpu.set_option("tokengen_exact", 0)
        for mode in (0, 2):
            gpu.set_option("tokengen_mode", mode)
            try:
                fast = gpu.tokengen_batch(base[s:s + 250_000])
                assert gpu.get_info("last_tokengen_path") == (3 if mode == 0 else 2)
            finally:
                gpu.set_option("tokengen_mode", 0)
            tot_re[mode] += gpu.get_info("tokengen_rechecked")
            assert gpu.get_info("tokengen_overflow") == 0
            assert np.array_equal(fast, exact), mode
    for mode in (0, 2):
        frac = tot_re[mode] / (base.shape[0] * 64 * 24)
        assert 0 < frac < 0.03, (mode, frac)                               # a small fraction needs the exact arithmetic


def test_search_matches_oracle_sample_and_is_invariant_under_rotate_migrate(c2):
    cfg, gpu, q = c2["cfg"], c2["gpu"], c2["queries"]
    k = cfg.k
    got = gpu.search_batch(q, k, cfg.probes, cfg.hard_cap, cfg.B)
    assert (got["n_ret"] == k).all() and (got["counters"][:, 5] == cfg.B).all()            # B candidates refined per query
    assert (got["counters"][:, 2] == cfg.B).all()                                           # all authenticated and decrypted
    assert (np.diff(got["top_dist"], axis=1) >= 0).all()                                    # sorted
    # oracle check on 400 queries: ids exact, FP64 distances bit-exact, counters equal
    g = O.GFunctions(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, c2["alpha"], c2["r"], c2["omega"])
    ix = O.Index(g, cfg.N, c2["mn"].shape[1], c2["mn"], c2["mx"], c2["rep"], c2["ids"])
    st = O.Store(cfg.dim, c2["iv"], c2["ct"], np.ones(cfg.N, dtype=np.int32), {1: c2["km"].derive(1)})
    NS = 400                                                                                # oracle sample (the C port does ~190 queries/s per core)
    codes = O.tokengen_batch(q[:NS], g)
    assert np.array_equal(gpu.tokengen_batch(q[:NS]), codes)
    for i in range(NS):
        ref = O.search(ix, st, q[i], codes[i], k, cfg.probes, cfg.hard_cap, cfg.B)
        assert np.array_equal(got["top_ids"][i], ref["top_ids"])
        assert np.array_equal(got["top_dist"][i].view(np.uint64), ref["top_dist"].view(np.uint64))
        c = got["counters"][i]
        assert (c[0], c[1], c[2], c[3], bool(c[4])) == (ref["cand_total"], ref["cand_kept"], ref["cand_decrypted"], ref["returned"], ref["retried"])
    # a full 10k-query batch takes the chunked, overlapped upload of the host entry (>= 4096 queries): same answers as 2000 at a time,
    # and as the plain single-copy upload
    q10 = WL.query_vectors(cfg, 10000)
    full = gpu.search_batch(q10, k, cfg.probes, cfg.hard_cap, cfg.B)
    for b0 in (0, 4000, 8000):
        part = gpu.search_batch(q10[b0:b0 + 2000], k, cfg.probes, cfg.hard_cap, cfg.B)
        for key in ("top_ids", "n_ret", "counters"):
            assert np.array_equal(part[key], full[key][b0:b0 + 2000]), key
        assert np.array_equal(part["top_dist"].view(np.uint64), full["top_dist"][b0:b0 + 2000].view(np.uint64))
    gpu.set_option("h2d_overlap", 0)
    plain = gpu.search_batch(q10, k, cfg.probes, cfg.hard_cap, cfg.B)
    gpu.set_option("h2d_overlap", 2)
    assert np.array_equal(plain["top_ids"], full["top_ids"]) and np.array_equal(plain["counters"], full["counters"])
    assert gpu.get_info("last_route_v2") == 1 and gpu.get_info("route_overflowed") == 0     # the two-CTA Route kernel held every query
    # config 5: Rotate -> v2, Migrate ids = 0 (mod 3) on the device with fresh IVs; then Rotate -> v3 and Migrate the touched set
    gpu.touched(clear=True)
    gpu.keys_set(2, c2["km"].derive(2))
    mig = np.arange(0, cfg.N, 3, dtype=np.int32)
    out = gpu.migrate(mig, WL.record_ivs(len(mig), 4242), 2)
    assert out["count"] == len(mig) and out["reencrypted"].all()
    assert not np.array_equal(out["ct"][:100], c2["ct"][mig[:100]])                         # ciphertexts really changed ...
    again = gpu.search_batch(q, k, cfg.probes, cfg.hard_cap, cfg.B)
    for key in ("top_ids", "n_ret", "counters"):                                            # ... results did not
        assert np.array_equal(again[key], got[key]), key
    assert np.array_equal(again["top_dist"].view(np.uint64), got["top_dist"].view(np.uint64))
    touched = gpu.touched(clear=True)
    assert 0 < len(touched) <= 2000 * cfg.B
    gpu.keys_set(3, c2["km"].derive(3))
    out3 = gpu.migrate(touched, WL.record_ivs(len(touched), 4243), 3)
    assert out3["count"] == len(touched)
    third = gpu.search_batch(q, k, cfg.probes, cfg.hard_cap, cfg.B)
    assert np.array_equal(third["top_ids"], got["top_ids"]) and np.array_equal(third["top_dist"].view(np.uint64), got["top_dist"].view(np.uint64))
    assert gpu.migrate(touched, WL.record_ivs(len(touched), 4244), 3)["count"] == 0         # idempotent
    # a migrated record decrypts under its new version with the oracle (OpenSSL): AAD and tag are the reference's
    j = 7
    rc, pt = O.decrypt_point(int(touched[j]), 3, cfg.dim, c2["km"].derive(3), out3["iv"][j].tobytes(), out3["ct"][j].tobytes())
    assert rc == 0 and np.array_equal(pt, c2["base"][touched[j]])


def test_groundtruth_and_recall_at_full_size(c2):
    cfg, gpu = c2["cfg"], c2["gpu"]
    q = c2["queries"][:256]
    b32, q32 = c2["base"].astype(np.float32), q.astype(np.float32)
    gt, d2 = gpu.groundtruth(b32, q32, 10, want_d2=True)
    assert (np.diff(d2, axis=1) >= 0).all()
    for i in (0, 100, 255):                                                                  # brute-force check of whole rows
        s = ((q32[i][None, :] - b32).astype(np.float64) ** 2).sum(1)                         # integer-valued data: exact in any order
        o = np.lexsort((np.arange(cfg.N), s))[:10]
        assert np.array_equal(gt[i], o) and np.array_equal(d2[i], s[o])
    res = gpu.search_batch(q, 10, cfg.probes, cfg.hard_cap, cfg.B)
    rec = gpu.recall_batch(gt, res["top_ids"], 10, res["n_ret"])
    assert 0.2 < rec.mean() <= 1.0
    # the true nearest neighbours found by the search carry the exact distance: sqrt of the ground-truth sum
    for i in range(256):
        hit = np.isin(res["top_ids"][i], gt[i])
        for pos in np.nonzero(hit)[0]:
            gi = int(np.nonzero(gt[i] == res["top_ids"][i][pos])[0][0])
            assert res["top_dist"][i][pos] == np.sqrt(d2[i][gi])


def test_config3_shape_at_full_size_matches_oracle_sample():
    """BASELINE configs[2] (GloVe-shape 1.2 M x 100, non-integer data: the FP64 distance path, 800-byte records) at full size: device
    Setup, a 10k-query batch, oracle sample with counters."""
    from fspann_query_system_b200.gpu import GpuContext
    cfg = WL.CONFIGS["C3"]
    base = WL.base_vectors(cfg)
    alpha, r, omega = HS.build_gfunctions(base[:HS.MIN_SAMPLE_SIZE], cfg.m, cfg.lam, cfg.seed, cfg.T, cfg.D)
    gpu = GpuContext(0)
    try:
        gpu.gfunctions_upload(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
        mn, mx, rep, ids = gpu.routing_build(base, HS.staged_order(cfg.N))
        km = HS.KeyManager(WL.MASTER_KEY)
        iv = WL.record_ivs(cfg.N, cfg.base_seed + 5)
        gpu.keys_set(1, km.derive(1))
        ct = gpu.encrypt_batch(np.arange(cfg.N, dtype=np.int32), base, iv, 1)
        gpu.store_upload(cfg.dim, iv, ct, np.ones(cfg.N, dtype=np.int32))
        q = WL.query_vectors(cfg, 10000)
        got = gpu.search_batch(q, cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
        assert (np.diff(got["top_dist"], axis=1)[got["n_ret"] == cfg.k] >= 0).all()
        g = O.GFunctions(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, alpha, r, omega)
        ix = O.Index(g, cfg.N, mn.shape[1], mn, mx, rep, ids)
        st = O.Store(cfg.dim, iv, ct, np.ones(cfg.N, dtype=np.int32), {1: km.derive(1)})
        sample = np.r_[0:100, 5000:5100, 9900:10000]
        codes = O.tokengen_batch(q[sample], g)
        for j, i in enumerate(sample):
            ref = O.search(ix, st, q[i], codes[j], cfg.k, cfg.probes, cfg.hard_cap, cfg.B)
            n = len(ref["top_ids"])
            assert got["n_ret"][i] == n and np.array_equal(got["top_ids"][i, :n], ref["top_ids"])
            assert np.array_equal(got["top_dist"][i, :n].view(np.uint64), ref["top_dist"].view(np.uint64))
            c = got["counters"][i]
            assert (c[0], c[1], c[2], c[3], bool(c[4])) == (ref["cand_total"], ref["cand_kept"], ref["cand_decrypted"], ref["returned"], ref["retried"])
    finally:
        gpu.close()


@pytest.mark.parametrize("probes,hard_cap,B", [(5, 24000, 1024), (5, 6000, 1024), (5, 12000, 4000), (6, 20000, 16000), (3, 2500, 1800), (5, 20000, 512),
                                               (5, 20416, 1024), (5, 20417, 2000), (7, 30000, 22000), (7, 20000, 20000), (5, 50000, 40000)])
def test_route_fast_path_with_binding_cap_and_long_lists_at_full_size(c2, probes, hard_cap, B):
    """The reference's own profiles (config_sift1m.json, sift1m_sub1.json, ...) use refinement limits of thousands and HARD_CAPs that bind.
    At 1 M ids the duplication is low, so the two-CTA Route kernel holds (nearly) every query itself: binding caps through its visit cut,
    B > 1024 through the separate sort kernel.  Ordered ids, scores and the raw / unique counters against the oracle."""
    cfg, gpu = c2["cfg"], c2["gpu"]
    NQ = 96
    q = c2["queries"][:NQ]
    g = O.GFunctions(cfg.dim, cfg.T, cfg.D, cfg.m, cfg.lam, c2["alpha"], c2["r"], c2["omega"])
    ix = O.Index(g, cfg.N, c2["mn"].shape[1], c2["mn"], c2["mx"], c2["rep"], c2["ids"])
    codes = O.tokengen_batch(q, g)
    gpu.set_option("route_small_v1", 0)                                  # 96 queries < SM count: keep them on the two-CTA kernel under test
    try:
        out = gpu.route_batch(codes, probes, hard_cap, B)
    finally:
        gpu.set_option("route_small_v1", 1)
    assert gpu.get_info("last_route_path") == 1 and gpu.get_info("last_route_v2") == 1
    assert gpu.get_info("route_overflowed") <= NQ // 10
    n_raw = cfg.T * cfg.D * probes * 64
    bound = 0
    for i in range(NQ):
        ids, sc, raw, mc = O.route(ix, codes[i], probes, hard_cap)
        n = min(B, len(ids))
        assert (out["n_cand"][i], out["unique"][i], out["raw_seen"][i]) == (n, len(ids), raw), i
        assert np.array_equal(out["cand_scores"][i, :n], sc[:n])
        if mc < 9:
            assert np.array_equal(out["cand_ids"][i, :n], ids[:n]), i
        bound += len(ids) >= hard_cap
    if hard_cap <= n_raw - 64 - 2000:
        assert bound == NQ                                               # the cap really cut every query's visits short
