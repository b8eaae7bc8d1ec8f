"""Randomised parity sweep: 14 small worlds with random shapes (N, dim, tables, divisions, m, lambda, data law) x random query
parameters (probes, hard cap, B, k) -- TokenGen codes, ordered candidate lists + counters (fast and general Route), full search
incl. the adaptive retry, verdicts with random tampering / deletions / retired key versions: everything bit-exact vs the oracle.
The seeds are fixed, so the sweep is deterministic; its job is to reach corners the hand-written cases do not (cut classes of
every size, worklist-heavy routes, keys that do not fit the 32-bit sort, ragged partitions, odd dimensions, W = 2 codes)."""
import numpy as np
import pytest

from fspann_query_system_b200 import _native as N
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SEEDS = list(range(14))


def random_world_kwargs(rng):
    lam = int(rng.integers(1, 5))
    m = int(rng.integers(4, 25))
    while m * lam > 126:
        m -= 1
    return dict(N=int(rng.integers(1000, 4200)), dim=int(rng.integers(3, 70)), Q=int(rng.integers(4, 20)), T=int(rng.integers(1, 5)),
                D=int(rng.integers(1, 6)), m=m, lam=lam, shape=str(rng.choice(["sift", "glove", "deep"])), data_seed=int(rng.integers(1, 10**6)),
                n_versions=int(rng.integers(1, 4)))


@pytest.mark.parametrize("seed", SEEDS)
def test_random_world_matches_oracle(seed, world_factory):
    rng = np.random.default_rng(1000 + seed)
    kw = random_world_kwargs(rng)
    w = world_factory(**kw)
    ctx = w.gpu_context()
    try:
        n, Q = w.cfg.N, w.queries.shape[0]
        codes = O.tokengen_batch(w.queries, w.g)
        assert np.array_equal(ctx.tokengen_batch(w.queries), codes), kw
        # ---- Route, 3 random parameter sets, both kernels
        for _ in range(3):
            probes = int(rng.integers(1, 12))
            n_raw = w.g.T * w.g.D * probes * 64
            hard_cap = int(rng.choice([rng.integers(1, 300), rng.integers(300, 3000), 20000, 1 << 20]))
            B = int(rng.choice([1, rng.integers(2, 64), rng.integers(64, 1100), rng.integers(1100, 3000)]))
            for general in (0, 1):
                ctx.set_option("route_general", general)
                try:
                    out = ctx.route_batch(codes, probes, hard_cap, B)
                finally:
                    ctx.set_option("route_general", 0)
                for q in range(Q):
                    ids, sc, raw, mc = O.route(w.ix, codes[q], probes, hard_cap)
                    k = min(B, len(ids))
                    assert (out["n_cand"][q], out["unique"][q], out["raw_seen"][q]) == (k, len(ids), raw), (kw, probes, hard_cap, B, general, q, n_raw)
                    assert np.array_equal(out["cand_scores"][q, :k], sc[:k])
                    if mc < 9:
                        assert np.array_equal(out["cand_ids"][q, :k], ids[:k]), (kw, probes, hard_cap, B, general, q)
        # ---- store with random damage: tampered tags, a deleted set, one key version retired
        store = O.Store(w.store.dim, w.store.iv.copy(), w.store.ct.copy(), w.store.key_version.copy(), dict(w.store.keys))
        for i in rng.choice(n, size=15, replace=False):
            store.ct[i, int(rng.integers(0, store.ct.shape[1]))] ^= int(rng.integers(1, 256))
        deleted = (rng.random(n) < 0.03).astype(np.uint8)
        if len(store.keys) > 1 and rng.random() < 0.7:
            dead = int(rng.choice(list(store.keys)))
            del store.keys[dead]
            ctx.keys_retire(dead)
        ctx.store_upload(w.g.dim, store.iv, store.ct, store.key_version)
        ctx.deleted_set(deleted)
        ix = O.Index(w.ix.g, w.ix.N, w.ix.P, w.ix.min_key, w.ix.max_key, w.ix.rep, w.ix.ids, deleted)
        st = O.Store(store.dim, store.iv, store.ct, store.key_version, store.keys, deleted)
        for _ in range(2):
            k = int(rng.choice([1, 5, 10, 40]))
            B = int(rng.choice([8, 64, 300, 1024]))
            probes = int(rng.integers(1, 8))
            hard_cap = int(rng.choice([2000, 20000]))
            ctx.touched(clear=True)
            got = ctx.search_batch(w.queries, k, probes, hard_cap, B)
            touched = np.zeros(n, dtype=np.uint8)
            for q in range(Q):
                ref = O.search(ix, st, w.queries[q], codes[q], k, probes, hard_cap, B, touched=touched)
                m = len(ref["top_ids"])
                assert got["n_ret"][q] == m, (kw, k, B, probes, q)
                assert np.array_equal(got["top_ids"][q, :m], ref["top_ids"])
                assert np.array_equal(got["top_dist"][q, :m].view(np.uint64), ref["top_dist"].view(np.uint64))
                c = got["counters"][q]
                assert (c[0], c[1], c[2], c[3], bool(c[4])) == (ref["cand_total"], ref["cand_kept"], ref["cand_decrypted"], ref["returned"], ref["retried"])
            assert np.array_equal(ctx.touched(clear=True), np.nonzero(touched)[0])
        # verdict classes on a direct refine call
        cand = np.tile(rng.integers(-2, n + 3, size=(1, 96)).astype(np.int32), (Q, 1))
        nc = np.full(Q, 96, dtype=np.int32)
        out = ctx.refine_batch(w.queries, cand, nc, 10)
        for q in range(0, Q, 3):
            ref = O.refine(st, w.queries[q], cand[q], 10)
            assert np.array_equal(out["verdict"][q, :96], ref["verdict"]) and out["n_decrypted"][q] == ref["n_decrypted"]
            assert np.array_equal(out["top_ids"][q, :len(ref["top_ids"])], ref["top_ids"])
    finally:
        ctx.close()
