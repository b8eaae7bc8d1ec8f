"""SURVEY 8(f)-1 / 8(f)-2 on the GPU: Migrate (KeyRotationServiceImpl.reencryptTouched, KRS:215-289) in place in the HBM store and
bulk encryptToPoint (AGC:55-112), both bit-exact against the oracle (OpenSSL AES-256-GCM) on the same inputs and IVs."""
import copy

import numpy as np
import pytest

from fspann_query_system_b200 import _native as N, workloads as WL
from fspann_query_system_b200.gpu import GpuContext
from oracle import oracle as O
from test_gpu_lifecycle import W1, W3, check_refine_verdicts, check_search

pytestmark = pytest.mark.gpu


def store_copy(st):
    return O.Store(st.dim, st.iv.copy(), st.ct.copy(), st.key_version.copy(), dict(st.keys))


@pytest.mark.parametrize("world", ["w3-dim32", "odd-dim33"])
def test_migrate_in_place_bit_exact_with_skips(world_factory, world):
    """Mixed stored versions 1..3, target 3 then 4: already-upgraded records, ids the store does not hold, duplicated list entries,
    a tampered record (tag failure) and records of a retired version are all skipped exactly like KRS:243-279; everything
    else gets the oracle's ciphertext and tag bit for bit, and the HBM store then answers searches like the oracle's store."""
    kw = dict(W3) if world == "w3-dim32" else dict(N=2000, dim=33, Q=24, T=2, D=3, m=10, lam=3, shape="glove", n_versions=3)
    w = world_factory(**kw)
    n = w.cfg.N
    st = store_copy(w.store)
    st.ct[17, 40] ^= 0x20                                               # tampered: decryptFromPoint throws -> skipped
    ctx = GpuContext(0)
    try:
        g, ix = w.g, w.ix
        ctx.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        for v, k in st.keys.items():
            ctx.keys_set(v, k)
        ctx.store_upload(g.dim, st.iv, st.ct, st.key_version)
        rng = np.random.default_rng(11)
        # --- pass 1: target = 3 with every kind of skip in the list
        ids = rng.permutation(n)[: n // 2].astype(np.int32)
        ids = np.concatenate([ids, ids[:50], np.array([17, -3, n + 9, 2**31 - 2], dtype=np.int32)])
        ivs = WL.record_ivs(len(ids), 9001)
        ref = store_copy(st)
        ver_before = ref.key_version.copy()
        n_ref = O.migrate(ref, np.where((ids >= 0) & (ids < n), ids, -1), ivs, 3)     # the oracle skips out-of-range ids the same way
        out = ctx.migrate(ids, ivs, 3)
        assert out["count"] == n_ref and out["reencrypted"].sum() == n_ref
        expect = np.zeros(len(ids), dtype=np.uint8)
        seen = set()
        for j, i in enumerate(ids.tolist()):
            if 0 <= i < n and i not in seen:
                seen.add(i)
                expect[j] = ver_before[i] < 3 and i != 17
        assert np.array_equal(out["reencrypted"], expect)
        sel = np.nonzero(expect)[0]
        assert np.array_equal(out["iv"][sel], ref.iv[ids[sel]]) and np.array_equal(out["iv"][sel], ivs[sel])
        assert np.array_equal(out["ct"][sel], ref.ct[ids[sel]]), "re-encrypted ciphertext||tag must equal OpenSSL's"
        assert (ref.key_version[ids[sel]] == 3).all()
        # the HBM store now equals the oracle's migrated store: same verdicts (the tampered record still fails), same results
        check_refine_verdicts(ctx, w, ref)
        check_search(ctx, w, ref)
        # --- pass 2: Rotate -> v4, Retire v1 while some records are still bound to it (a host would refuse; the device must cope):
        # v1 records are skipped (no key), v2/v3 records move to v4
        k4 = O.kdf(w.master, 4)
        ctx.keys_set(4, k4); ref.keys[4] = k4
        ctx.keys_retire(1); del ref.keys[1]
        ids2 = np.arange(n, dtype=np.int32)
        ivs2 = WL.record_ivs(n, 9002)
        ver2 = ref.key_version.copy()
        n_ref2 = O.migrate(ref, ids2, ivs2, 4)
        out2 = ctx.migrate(ids2, ivs2, 4)
        assert out2["count"] == n_ref2
        assert np.array_equal(out2["reencrypted"].astype(bool), (ver2 > 1) & (np.arange(n) != 17))
        sel2 = np.nonzero(out2["reencrypted"])[0]
        assert np.array_equal(out2["ct"][sel2], ref.ct[sel2]) and np.array_equal(out2["iv"][sel2], ref.iv[sel2])
        assert N.V_NO_KEY in check_refine_verdicts(ctx, w, ref)          # records still bound to the retired v1
        check_search(ctx, w, ref)
        # idempotent: nothing is left below the target
        assert ctx.migrate(ids2, WL.record_ivs(n, 9003), 4)["count"] == 0
    finally:
        ctx.close()


def test_migrate_argument_errors(world_factory):
    w = world_factory(**W1)
    ctx = w.gpu_context()
    try:
        with pytest.raises(N.IllegalArgumentError):
            ctx.migrate(np.arange(4, dtype=np.int32), WL.record_ivs(4, 1), 7)          # target version without a key
        assert ctx.migrate(np.zeros(0, dtype=np.int32), np.zeros((0, 12), dtype=np.uint8), 1)["count"] == 0   # KRS:220-223
        fresh = GpuContext(0)
        try:
            fresh.store_dim = 32
            with pytest.raises(N.IllegalStateError):
                fresh.migrate(np.arange(4, dtype=np.int32), WL.record_ivs(4, 1), 1)    # no store
        finally:
            fresh.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("dim", [32, 33, 100, 128])
def test_encrypt_batch_bit_exact(dim):
    """Bulk encryptToPoint: ciphertext || tag equals OpenSSL's for arbitrary (non-contiguous) ids, incl. odd dimensions,
    signed zeros, subnormals and the largest finite values; works before any store exists."""
    rng = np.random.default_rng(dim)
    n = 700
    vec = rng.normal(0, 50, size=(n, dim))
    vec[0, :4] = [0.0, -0.0, 5e-324, np.finfo(np.float64).max]
    ids = rng.choice(2_000_000, size=n, replace=False).astype(np.int32)
    ivs = WL.record_ivs(n, 77 + dim)
    key = O.kdf(WL.MASTER_KEY, 5)
    ctx = GpuContext(0)
    try:
        ctx.keys_set(5, key)
        ct = ctx.encrypt_batch(ids, vec, ivs, 5)
        assert np.array_equal(ct, O.encrypt_store(vec, 5, key, ivs, ids=ids))
        rc, back = O.decrypt_point(int(ids[3]), 5, dim, key, ivs[3].tobytes(), ct[3].tobytes())
        assert rc == 0 and np.array_equal(back.view(np.uint64), vec[3].view(np.uint64))
        with pytest.raises(N.IllegalArgumentError):
            ctx.encrypt_batch(ids, vec, ivs, 6)                                        # no such key version
    finally:
        ctx.close()


def test_setup_encrypt_on_device_then_search(world_factory):
    """Setup with the device doing the encryption: encrypt_batch -> store_upload -> search equals the oracle's."""
    w = world_factory(**W1)
    ctx = GpuContext(0)
    try:
        g, ix = w.g, w.ix
        ctx.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        ctx.keys_set(1, w.keys[1])
        ct = ctx.encrypt_batch(np.arange(w.cfg.N, dtype=np.int32), w.base, w.iv, 1)
        assert np.array_equal(ct, w.ct)
        ctx.store_upload(g.dim, w.iv, ct, w.key_version)
        check_search(ctx, w, w.store)
    finally:
        ctx.close()
