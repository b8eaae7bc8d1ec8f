"""Multi-GPU logic: world_size-2 gloo run on CPU (collective + merge + retry agreement, with an oracle-backed shard
engine), and a single-GPU run with two store shards in two contexts (the sharded CUDA path end to end)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fspann_query_system_b200 import distributed as DD
from oracle import oracle as O


def test_split_and_shard_ranges_cover_everything():
    for n in (0, 1, 7, 10000, 1_200_000):
        for world in (1, 2, 3, 8):
            parts = [DD.split_batch(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def test_merge_topk_is_a_stable_merge():
    rng = np.random.default_rng(0)
    S, Q, k = 3, 50, 7
    # ground truth: one list per query with many exact ties, split over shards by candidate ownership
    d = np.full((S, Q, k), np.nan); r = np.full((S, Q, k), DD.RANK_PAD, np.int32); i = np.full((S, Q, k), -1, np.int32)
    want_ids = np.full((Q, k), -1, np.int32)
    for q in range(Q):
        n = rng.integers(0, 40)
        dist_all = rng.integers(0, 6, size=n).astype(np.float64)      # heavy ties
        owner = rng.integers(0, S, size=n)
        order = np.lexsort((np.arange(n), dist_all))[:k]
        want_ids[q, :len(order)] = order + 1000
        for s in range(S):
            mine = np.nonzero(owner == s)[0]
            loc = mine[np.lexsort((mine, dist_all[mine]))][:k]
            d[s, q, :len(loc)], r[s, q, :len(loc)], i[s, q, :len(loc)] = dist_all[loc], loc, loc + 1000
    ids, dd, nret = DD.merge_topk(torch.from_numpy(d), torch.from_numpy(r), torch.from_numpy(i), k)
    assert np.array_equal(ids.numpy(), want_ids)
    assert np.array_equal(nret.numpy(), (want_ids >= 0).sum(1))


class OracleShardEngine:
    """CPU stand-in for a GpuContext holding the store shard [lo, hi): same method surface, oracle arithmetic."""

    def __init__(self, w, lo, hi):
        self.w, self.lo, self.hi = w, lo, hi
        present = np.zeros(w.cfg.N, dtype=np.uint8)
        present[lo:hi] = 1
        self.store = O.Store(w.store.dim, w.iv, w.ct, w.key_version, dict(w.store.keys), None, present)

    def tokengen_batch(self, queries):
        return O.tokengen_batch(queries, self.w.g)

    def route_batch(self, codes, probes, hard_cap, B):
        Q = codes.shape[0]
        cid = np.full((Q, B), -1, np.int32); nc = np.zeros(Q, np.int32); raw = np.zeros(Q, np.int32); uq = np.zeros(Q, np.int32)
        for q in range(Q):
            ids, _, r, _ = O.route(self.w.ix, codes[q], probes, hard_cap)
            n = min(B, len(ids)); cid[q, :n] = ids[:n]; nc[q] = n; raw[q] = r; uq[q] = len(ids)
        return dict(cand_ids=cid, n_cand=nc, raw_seen=raw, unique=uq)

    def refine_batch(self, queries, cand, ncand, k):
        Q = cand.shape[0]
        tid = np.full((Q, k), -1, np.int32); td = np.full((Q, k), np.nan); tr = np.full((Q, k), DD.RANK_PAD, np.int32)
        nd = np.zeros(Q, np.int32)
        for q in range(Q):
            c = cand[q, :ncand[q]]
            ref = O.refine(self.store, queries[q], c, k)
            ok = np.nonzero(ref["verdict"] == 0)[0]
            order = ok[np.lexsort((ok, ref["cand_dist"][ok]))][:k]
            tid[q, :len(order)], td[q, :len(order)], tr[q, :len(order)] = c[order], ref["cand_dist"][order], order
            nd[q] = len(ok)
        return dict(top_ids=tid, top_dist=td, top_rank=tr, n_decrypted=nd)


def _gloo_worker(rank, world, port, kw, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import World
        w = World(**kw)
        lo, hi = DD.shard_range(w.cfg.N, rank, world)
        s = DD.ShardedSearcher(OracleShardEngine(w, lo, hi), "cpu")
        res = {}
        for (k, B) in [(10, 128), (50, 100)]:           # the second forces the adaptive retry on every query
            res[(k, B)] = s.search_batch(w.queries, k, 5, 20000, B)
        lo_q, hi_q = DD.split_batch(w.queries.shape[0], rank, world)
        out.put((rank, res, (lo_q, hi_q)))
    finally:
        dist.destroy_process_group()


def test_sharded_search_two_ranks_gloo(world_factory):
    kw = dict(N=2000, dim=16, Q=12, T=2, D=3, m=12, lam=2)
    w = world_factory(**kw)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, kw, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([out.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    codes = O.tokengen_batch(w.queries, w.g)
    for (k, B) in [(10, 128), (50, 100)]:
        a, b = got[0][1][(k, B)], got[1][1][(k, B)]
        for key in ("top_ids", "n_ret", "n_decrypted", "retried"):
            assert np.array_equal(a[key], b[key]), "ranks must agree after the collective"
        for q in range(w.queries.shape[0]):
            ref = O.search(w.ix, w.store, w.queries[q], codes[q], k, 5, 20000, B)
            n = len(ref["top_ids"])
            assert a["n_ret"][q] == n and np.array_equal(a["top_ids"][q, :n], ref["top_ids"])
            assert np.array_equal(a["top_dist"][q, :n].view(np.uint64), ref["top_dist"].view(np.uint64))
            assert a["n_decrypted"][q] == ref["cand_decrypted"] and bool(a["retried"][q]) == ref["retried"]
    assert got[0][2] == (0, 6) and got[1][2] == (6, 12)


@pytest.mark.gpu
def test_sharded_store_on_gpu_matches_single_store(world_factory):
    """Two store shards in two contexts on one GPU: per-shard refine (global ids in the AAD, OTHER_SHARD verdicts), merge on
    (distance, rank) == unsharded search == oracle."""
    from fspann_query_system_b200 import _native as N
    from fspann_query_system_b200.gpu import GpuContext
    w = world_factory(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2, n_versions=3)
    g, ix = w.g, w.ix
    shards = []
    for r in range(2):
        lo, hi = DD.shard_range(w.cfg.N, r, 2)
        c = GpuContext(0)
        c.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        for v, key in w.store.keys.items():
            c.keys_set(v, key)
        c.store_upload(g.dim, w.iv[lo:hi], w.ct[lo:hi], w.key_version[lo:hi], id_base=lo, n_global=w.cfg.N)
        shards.append((c, lo, hi))
    try:
        k, B = 10, 128            # >= 10*k decrypted, so the reference does not take its adaptive retry
        codes = O.tokengen_batch(w.queries, w.g)
        route = shards[0][0].route_batch(codes, 5, 20000, B)
        outs = [c.refine_batch(w.queries, route["cand_ids"], route["n_cand"], k) for c, _, _ in shards]
        for (c, lo, hi), o in zip(shards, outs):
            for q in range(w.queries.shape[0]):
                cand = route["cand_ids"][q, :route["n_cand"][q]]
                mine = (cand >= lo) & (cand < hi)
                assert np.all(o["verdict"][q, :len(cand)][~mine] == N.V_OTHER_SHARD)
                assert np.all(o["verdict"][q, :len(cand)][mine] == N.V_OK)
            t = c.touched(clear=True)
            assert len(t) > 0 and t.min() >= lo and t.max() < hi
        dv = torch.stack([torch.from_numpy(np.where(o["top_ids"] >= 0, o["top_dist"], np.inf)) for o in outs])
        ids, dd, nret = DD.merge_topk(dv, torch.stack([torch.from_numpy(o["top_rank"]) for o in outs]),
                                      torch.stack([torch.from_numpy(o["top_ids"]) for o in outs]), k)
        for q in range(w.queries.shape[0]):
            ref = O.search(w.ix, w.store, w.queries[q], codes[q], k, 5, 20000, B)
            assert np.array_equal(ids[q].numpy(), ref["top_ids"])
            assert np.array_equal(dd[q].numpy().view(np.uint64), ref["top_dist"].view(np.uint64))
            assert outs[0]["n_decrypted"][q] + outs[1]["n_decrypted"][q] == ref["cand_decrypted"]
        # store_update addresses records by GLOBAL id and refuses ids of another shard
        c0, lo0, hi0 = shards[0]
        with pytest.raises(N.IllegalArgumentError):
            c0.store_update(np.array([hi0 + 1], np.int32), w.iv[:1], w.ct[:1], w.key_version[:1])
        s1 = DD.ShardedSearcher(shards[1][0], "cpu")          # world of one: degenerates to the local shard
        assert s1.search_batch(w.queries[:4], k, 5, 20000, B)["top_ids"].shape == (4, k)
    finally:
        for c, _, _ in shards:
            c.close()


@pytest.mark.gpu
def test_device_resident_sharded_blocks_and_merge_kernel(world_factory):
    """Device-resident config-4 building blocks on one GPU: route_batch_dev + refine_batch_dev on two store shards (two contexts) +
    merge_topk_kernel == unsharded search == oracle, incl. retried queries; DeviceShardedSearcher with a world of one."""
    from fspann_query_system_b200.gpu import GpuContext
    w = world_factory(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2, n_versions=3)
    g, ix = w.g, w.ix
    dev = torch.device("cuda", 0)
    shards = []
    for r in range(2):
        lo, hi = DD.shard_range(w.cfg.N, r, 2)
        c = GpuContext(0)
        c.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        for v, key in w.store.keys.items():
            c.keys_set(v, key)
        c.store_upload(g.dim, w.iv[lo:hi], w.ct[lo:hi], w.key_version[lo:hi], id_base=lo, n_global=w.cfg.N)
        shards.append(c)
    full = w.gpu_context()
    try:
        Q, k, B = w.queries.shape[0], 10, 128
        dq = torch.from_numpy(w.queries).to(dev)
        i32 = dict(dtype=torch.int32, device=dev)
        cand = torch.full((Q, B), -1, **i32); ncand = torch.zeros(Q, **i32)
        shards[0].route_batch_dev(Q, dq.data_ptr(), 5, 20000, B, cand.data_ptr(), ncand.data_ptr())
        shards[0].sync()
        codes = O.tokengen_batch(w.queries, w.g)
        for q in range(Q):
            rids = O.route(w.ix, codes[q], 5, 20000)[0]
            n = min(B, len(rids))
            assert int(ncand[q]) == n and np.array_equal(cand[q, :n].cpu().numpy(), rids[:n])
        ids = torch.empty((2, Q, k), **i32); rk = torch.empty((2, Q, k), **i32)
        dd = torch.empty((2, Q, k), dtype=torch.float64, device=dev)
        nret = torch.empty((2, Q), **i32); ndec = torch.empty((2, Q), **i32)
        for s, c in enumerate(shards):
            c.refine_batch_dev(Q, dq.data_ptr(), cand.data_ptr(), ncand.data_ptr(), B, k, ids[s].data_ptr(), dd[s].data_ptr(), rk[s].data_ptr(),
                               nret[s].data_ptr(), ndec[s].data_ptr())
            c.sync()
        out_i = torch.empty((Q, k), **i32); out_d = torch.empty((Q, k), dtype=torch.float64, device=dev); out_n = torch.empty(Q, **i32)
        shards[0].merge_topk_dev(2, Q, k, dd.data_ptr(), rk.data_ptr(), ids.data_ptr(), out_i.data_ptr(), out_d.data_ptr(), out_n.data_ptr())
        shards[0].sync()
        for q in range(Q):
            ref = O.search(w.ix, w.store, w.queries[q], codes[q], k, 5, 20000, B)
            assert np.array_equal(out_i[q].cpu().numpy(), ref["top_ids"]) and int(out_n[q]) == len(ref["top_ids"])
            assert np.array_equal(out_d[q].cpu().numpy().view(np.uint64), ref["top_dist"].view(np.uint64))
            assert int(ndec[0, q] + ndec[1, q]) == ref["cand_decrypted"]
        # a world of one over the unsharded store, with the adaptive retry (k=100, B=256 forces decrypted < 10*K)
        s1 = DD.DeviceShardedSearcher(full)
        for kk, BB in [(10, 64), (100, 256)]:
            got = s1.search_batch_dev(dq, kk, 5, 20000, BB)
            torch.cuda.synchronize()
            n_retried = 0
            for q in range(Q):
                ref = O.search(w.ix, w.store, w.queries[q], codes[q], kk, 5, 20000, BB)
                n = len(ref["top_ids"])
                assert int(got["n_ret"][q]) == n and np.array_equal(got["top_ids"][q, :n].cpu().numpy(), ref["top_ids"])
                assert np.array_equal(got["top_dist"][q, :n].cpu().numpy().view(np.uint64), ref["top_dist"].view(np.uint64))
                assert bool(got["retried"][q]) == ref["retried"] and int(got["n_decrypted"][q]) == ref["cand_decrypted"]
                assert (int(got["raw_seen"][q]), int(got["unique"][q])) == (ref["cand_total"], ref["cand_kept"])
                n_retried += ref["retried"]
            if kk == 100:
                assert n_retried > 0
    finally:
        for c in shards:
            c.close()
        full.close()


# ---- database-sharded search as ONE C-ABI call (NCCL inside libfspann_gpu.so) -------------------------------------------------------
def _upload_shard(w, dev, lo, hi, n_ranks, rank, comm_id):
    from fspann_query_system_b200.gpu import GpuContext
    g, ix = w.g, w.ix
    c = GpuContext(dev)
    c.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
    for v, key in w.store.keys.items():
        c.keys_set(v, key)
    c.store_upload(g.dim, w.store.iv[lo:hi], w.store.ct[lo:hi], w.store.key_version[lo:hi], id_base=lo, n_global=w.cfg.N)
    c.comm_init(n_ranks, rank, comm_id)
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("k,B", [(10, 256), (100, 256)])
def test_sharded_abi_single_rank_equals_unsharded(world_factory, k, B):
    """fspann_sharded_search_batch with a one-rank communicator (no NCCL needed) walks the whole sharded pipeline -- slice routing,
    refine with candidate ranks, merge, device-side retry (k=100, B=256 retries every query) -- and must equal fspann_search_batch."""
    w = world_factory(N=20000, dim=128, Q=64, T=4, D=8, m=24, lam=2)
    full = w.gpu_context()
    sh = _upload_shard(w, 0, 0, w.cfg.N, 1, 0, None)
    try:
        ref = full.search_batch(w.queries, k, 5, 20000, B)
        got = sh.sharded_search_batch(w.queries, k, 5, 20000, B)
        for key in ("top_ids", "n_ret", "counters"):
            assert np.array_equal(got[key], ref[key]), key
        assert np.array_equal(got["top_dist"].view(np.uint64), ref["top_dist"].view(np.uint64))
        assert (got["counters"][:, 4] == (1 if k == 100 else 0)).all()
        st = sh.sharded_stage_ms()
        assert st["launches"] > 0 and st["gather_bytes"] == 0
        bad = w.queries.copy(); bad[5, 7] = np.inf
        with pytest.raises(Exception):
            sh.sharded_search_batch(bad, k, 5, 20000, B)
    finally:
        full.close(); sh.close()


@pytest.mark.gpu
def test_sharded_abi_two_gpus_one_process_threads(world_factory):
    """Two contexts on two GPUs in ONE process (a thread per GPU, the JVM deployment of INTEGRATION.md): comm_init + sharded search
    from both threads; both ranks return the unsharded result bit for bit, with mixed key versions and the retry pass."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w = world_factory(N=20000, dim=128, Q=64, T=4, D=8, m=24, lam=2, n_versions=2)
    full = w.gpu_context()
    cid = full.comm_unique_id()
    half = w.cfg.N // 2 + 37
    outs, errs, ctxs = {}, [], [None, None]

    def run(rank):
        try:
            lo, hi = (0, half) if rank == 0 else (half, w.cfg.N)
            c = _upload_shard(w, rank, lo, hi, 2, rank, cid)
            ctxs[rank] = c
            outs[rank] = [c.sharded_search_batch(w.queries, k, 5, 20000, 256) for k in (10, 100)]
            outs[(rank, "stage")] = c.sharded_stage_ms()
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in ts]; [t.join(timeout=300) for t in ts]
    try:
        assert not errs, errs
        for i, k in enumerate((10, 100)):
            ref = full.search_batch(w.queries, k, 5, 20000, 256)
            for rank in range(2):
                got = outs[rank][i]
                for key in ("top_ids", "n_ret", "counters"):
                    assert np.array_equal(got[key], ref[key]), (rank, k, key)
                assert np.array_equal(got["top_dist"].view(np.uint64), ref["top_dist"].view(np.uint64))
        assert outs[(0, "stage")]["gather_bytes"] > 0
    finally:
        full.close()
        for c in ctxs:
            if c is not None:
                c.close()
