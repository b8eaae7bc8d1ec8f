"""Multi-GPU forms of the hot path, one process per GPU over torch.distributed.

* query-parallel (BASELINE config 3): queries are independent (QSI.search shares no mutable state across queries), so the
  batch is split contiguously over ranks, every rank holds a replica of the routing index + store, and there is NO
  collective on the data path (`split_batch`; bench.py uses it for its weak-scaling runs).
* database-sharded (BASELINE config 4), device-resident form `DeviceShardedSearcher`: Route is query-parallel on the replicated
  routing index (each rank routes its slice of the batch), the candidate lists are all-gathered (Q*B*4 bytes), every rank refines
  for ALL queries the candidates its store shard holds, the per-shard top-k (distance, candidate rank, id) is all-gathered and merged
  by `merge_topk_kernel` on (distance, rank).  Everything stays in HBM; the two all-gathers are NCCL over NVLink.
* database-sharded, host-buffer form `ShardedSearcher` (also what the gloo CPU tests drive): the routing index is replicated, so
  every rank derives the IDENTICAL ordered candidate list; the encrypted store is sharded by contiguous global-id range; each rank authenticates + decrypts + scores
  only the candidates it owns and emits its local top-k as (distance, candidate rank, id); ONE all-gather (NCCL over
  NVLink on GPUs, gloo in the CPU tests) of Q*k*(8+4+4) bytes per rank, then a k-way merge keyed on (distance, rank) --
  the rank reproduces the reference's stable sort tie order (QSI:298), so the merged result is bit-identical to a
  single-store search.  The adaptive retry (QSI:327-337) uses the all-reduced decrypted count.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

RANK_PAD = 0x7FFFFFFF


def split_batch(Q: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of a Q-query batch owned by `rank`."""
    base, rem = divmod(Q, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_range(N: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous global-id range [lo, hi) of the store shard owned by `rank`."""
    return split_batch(N, rank, world)


def merge_topk(dists: torch.Tensor, ranks: torch.Tensor, ids: torch.Tensor, k: int):
    """dists f64 / ranks i32 / ids i32 of shape [S, Q, k] (per-shard results, padded with id=-1, rank=RANK_PAD).
    Returns (ids [Q,k], dists [Q,k], n_ret [Q]) ordered by (distance, candidate rank)."""
    S, Q, kk = dists.shape
    d = dists.permute(1, 0, 2).reshape(Q, S * kk).clone()
    r = ranks.permute(1, 0, 2).reshape(Q, S * kk)
    i = ids.permute(1, 0, 2).reshape(Q, S * kk)
    valid = i >= 0
    d[~valid] = float("inf")
    # lexicographic (d, r): stable sort by r, then stable sort by d
    o1 = torch.sort(r, dim=1, stable=True).indices
    d1 = torch.gather(d, 1, o1)
    o2 = torch.sort(d1, dim=1, stable=True).indices
    order = torch.gather(o1, 1, o2)[:, :k]
    out_i = torch.gather(i, 1, order)
    out_d = torch.gather(d, 1, order)
    n_ret = torch.clamp(valid.sum(dim=1), max=k).to(torch.int32)
    pad = torch.arange(k, device=d.device)[None, :] >= n_ret[:, None]
    out_i = torch.where(pad, torch.full_like(out_i, -1), out_i)
    out_d = torch.where(pad, torch.full_like(out_d, float("nan")), out_d)
    return out_i, out_d, n_ret


class ShardedSearcher:
    """search_batch over a store sharded across the ranks of `group`.  `engine` is a GpuContext holding the replicated routing
    index and this rank's store shard (any object with tokengen_batch / route_batch / refine_batch works: the CPU tests plug
    in an oracle-backed stand-in to exercise the collective logic under gloo)."""

    def __init__(self, engine, device: torch.device | str = "cpu", group=None):
        self.engine, self.device, self.group = engine, torch.device(device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t[None]
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return torch.stack(out)

    def _pass(self, queries, k, probes, hard_cap, B):
        codes = self.engine.tokengen_batch(queries)
        r = self.engine.route_batch(codes, probes, hard_cap, B)
        f = self.engine.refine_batch(queries, r["cand_ids"], r["n_cand"], k)
        dv = torch.from_numpy(np.where(f["top_ids"] >= 0, f["top_dist"], np.inf)).to(self.device)
        rk = torch.from_numpy(f["top_rank"]).to(self.device)
        iv = torch.from_numpy(f["top_ids"]).to(self.device)
        nd = torch.from_numpy(f["n_decrypted"].astype(np.int64)).to(self.device)
        ids, dd, nret = merge_topk(self._gather(dv), self._gather(rk), self._gather(iv), k)
        if self.world > 1:
            dist.all_reduce(nd, group=self.group)
        return ids, dd, nret, nd, r

    def search_batch(self, queries: np.ndarray, k: int, probes: int, hard_cap: int, B: int):
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        ids, dd, nret, nd, r = self._pass(queries, k, probes, hard_cap, B)
        retried = torch.zeros_like(nret, dtype=torch.bool)
        need = (nd > 0) & ((nret < k) | (nd < 10 * k))             # QSI:293 (no retry when nothing decrypted), QSI:444-447
        rows = torch.nonzero(need).flatten()
        if rows.numel() > 0:                                        # identical on every rank: all inputs were all-reduced
            sel = rows.cpu().numpy()
            ids2, dd2, nret2, nd2, _ = self._pass(queries[sel], k, 10, hard_cap, B)
            ids[rows], dd[rows], nret[rows], nd[rows] = ids2, dd2, nret2, nd2
            retried[rows] = True
        return dict(top_ids=ids.cpu().numpy(), top_dist=dd.cpu().numpy(), n_ret=nret.cpu().numpy(), n_decrypted=nd.cpu().numpy(),
                    retried=retried.cpu().numpy(), unique=r["unique"], raw_seen=r["raw_seen"])


class DeviceShardedSearcher:
    """Database-sharded search with everything resident in HBM (BASELINE config 4).  `gpu` holds the replicated routing index and
    this rank's store shard (store_upload(..., id_base=lo, n_global=N)).  Queries are a CUDA FP64 tensor [Q, dim] replicated on every
    rank.  world_size 1 (no process group) degenerates to a single shard and is what the single-GPU tests exercise."""

    def __init__(self, gpu, group=None):
        self.gpu, self.group = gpu, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = torch.device("cuda", gpu.device)
        self.stream = torch.cuda.ExternalStream(gpu.stream(), device=self.dev)

    def _all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """[...] -> [world, ...] (NCCL all-gather on the library's stream)."""
        if self.world == 1:
            return t[None]
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def _pass(self, dq: torch.Tensor, k: int, probes: int, hard_cap: int, B: int):
        Q = dq.shape[0]
        g, W = self.gpu, self.world
        per = (Q + W - 1) // W                                   # rows per rank (the last ranks may own fewer; padded for the gather)
        lo, hi = min(Q, self.rank * per), min(Q, (self.rank + 1) * per)
        i32 = dict(dtype=torch.int32, device=self.dev)
        cand = torch.full((per, B), -1, **i32)
        meta = torch.zeros((3, per), **i32)                      # n_cand, raw_seen, unique of my slice
        if hi > lo:
            g.route_batch_dev(hi - lo, dq[lo:hi].data_ptr(), probes, hard_cap, B, cand.data_ptr(), meta[0].data_ptr(), meta[1].data_ptr(), meta[2].data_ptr())
        cand_all = self._all_gather(cand).reshape(W * per, B)[:Q]                 # every rank: the identical ordered candidate lists
        meta_all = self._all_gather(meta).permute(1, 0, 2).reshape(3, W * per)[:, :Q].contiguous()
        ids = torch.empty((Q, k), **i32)
        rk = torch.empty((Q, k), **i32)
        dd = torch.empty((Q, k), dtype=torch.float64, device=self.dev)
        cnt = torch.empty((2, Q), **i32)                         # n_ret, n_decrypted on this shard
        cand_all = cand_all.contiguous()
        g.refine_batch_dev(Q, dq.data_ptr(), cand_all.data_ptr(), meta_all[0].data_ptr(), B, k, ids.data_ptr(), dd.data_ptr(), rk.data_ptr(),
                           cnt[0].data_ptr(), cnt[1].data_ptr())
        nd = cnt[1].to(torch.int64)
        if W > 1:
            dist.all_reduce(nd, group=self.group)
        out_i = torch.empty((Q, k), **i32)
        out_d = torch.empty((Q, k), dtype=torch.float64, device=self.dev)
        out_n = torch.empty((Q,), **i32)
        a_d, a_r, a_i = self._all_gather(dd), self._all_gather(rk), self._all_gather(ids)
        g.merge_topk_dev(W, Q, k, a_d.data_ptr(), a_r.data_ptr(), a_i.data_ptr(), out_i.data_ptr(), out_d.data_ptr(), out_n.data_ptr())
        return out_i, out_d, out_n, nd, meta_all

    def search_batch_dev(self, dq: torch.Tensor, k: int, probes: int, hard_cap: int, B: int, allow_retry: bool = True):
        """Returns device tensors dict(top_ids [Q,k], top_dist [Q,k], n_ret [Q], n_decrypted [Q], retried [Q], raw_seen, unique)."""
        assert dq.is_cuda and dq.dtype == torch.float64 and dq.is_contiguous()
        with torch.cuda.stream(self.stream):
            ids, dd, nret, nd, meta = self._pass(dq, k, probes, hard_cap, B)
            retried = torch.zeros_like(nret, dtype=torch.bool)
            if allow_retry:
                need = (nd > 0) & ((nret < k) | (nd < 10 * k))         # QSI:293, QSI:444-447; identical on every rank
                rows = torch.nonzero(need).flatten()                   # (synchronises: the retry decision is made on the host side)
                if rows.numel() > 0:
                    ids2, dd2, nret2, nd2, meta2 = self._pass(dq[rows].contiguous(), k, 10, hard_cap, B)
                    ids[rows], dd[rows], nret[rows], nd[rows] = ids2, dd2, nret2, nd2
                    meta[:, rows] = meta2
                    retried[rows] = True
        return dict(top_ids=ids, top_dist=dd, n_ret=nret, n_decrypted=nd, retried=retried, raw_seen=meta[1], unique=meta[2])
