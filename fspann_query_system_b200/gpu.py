"""Thin object wrapper over the C ABI: one GpuContext per GPU (include/fspann_gpu.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


class GpuContext:
    def __init__(self, device: int = 0, debug: bool = False):
        self.lib = N.load(debug)
        self.ctx = C.c_void_p()
        rc = self.lib.fspann_ctx_create(C.c_int(device), C.byref(self.ctx))
        if rc != N.OK:
            self.ctx = None
            N.check(self.lib, None, rc)
        self.device = device
        self.dim = self.T = self.D = self.m = self.lam = self.W = 0
        self.N = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.fspann_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        N.check(self.lib, self.ctx, rc)

    # ---- uploads
    def gfunctions_upload(self, dim, T, D, m, lam, alpha, r, omega):
        alpha, r, omega = (np.ascontiguousarray(x, dtype=np.float64) for x in (alpha, r, omega))
        self._ck(self.lib.fspann_gfunctions_upload(self.ctx, dim, T, D, m, lam, N.ptr(alpha), N.ptr(r), N.ptr(omega)))
        self.dim, self.T, self.D, self.m, self.lam, self.W = dim, T, D, m, lam, (m * lam + 63) // 64

    def routing_upload(self, dim, T, D, m, lam, alpha, r, omega, min_key, max_key, rep, ids):
        alpha, r, omega = (np.ascontiguousarray(x, dtype=np.float64) for x in (alpha, r, omega))
        min_key = np.ascontiguousarray(min_key, dtype=np.int64)
        max_key = np.ascontiguousarray(max_key, dtype=np.int64)
        rep = np.ascontiguousarray(rep, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n_ids = ids.shape[-1]
        self._ck(self.lib.fspann_routing_upload(self.ctx, dim, T, D, m, lam, N.ptr(alpha), N.ptr(r), N.ptr(omega), C.c_int64(n_ids),
                                                N.ptr(min_key), N.ptr(max_key), N.ptr(rep), N.ptr(ids)))
        self.dim, self.T, self.D, self.m, self.lam, self.W = dim, T, D, m, lam, (m * lam + 63) // 64

    def routing_build(self, vectors, staged_ids, want_arrays: bool = True):
        """finalizeForSearch on the device: code the base set + GreedyPartitioner.build per (t,d); installs the routing state.
        Returns (min_key, max_key, rep, ids) in routing_upload's layout when want_arrays."""
        vectors = np.ascontiguousarray(vectors, dtype=np.float64)
        staged_ids = np.ascontiguousarray(staged_ids, dtype=np.int32)
        n = vectors.shape[0]
        if vectors.ndim != 2 or vectors.shape[1] != self.dim:
            raise N.IllegalArgumentError(f"Expected vector length {self.dim}")
        TD, P = self.T * self.D, (n + 63) // 64
        mn = mx = rep = ids = None
        if want_arrays:
            mn, mx = np.empty((TD, P), dtype=np.int64), np.empty((TD, P), dtype=np.int64)
            rep, ids = np.empty((TD, P, self.W), dtype=np.uint64), np.empty((TD, n), dtype=np.int32)
        self._ck(self.lib.fspann_routing_build(self.ctx, C.c_int64(n), N.ptr(vectors), N.ptr(staged_ids), N.ptr(mn), N.ptr(mx), N.ptr(rep),
                                               N.ptr(ids)))
        return mn, mx, rep, ids

    # ---- the same build in pieces (a base set that does not fit the host: BASELINE config 4)
    def routing_build_begin(self, n: int):
        self._ck(self.lib.fspann_routing_build_begin(self.ctx, C.c_int64(n)))

    def routing_build_add(self, first_id: int, vectors):
        vectors = np.ascontiguousarray(vectors, dtype=np.float64)
        self._ck(self.lib.fspann_routing_build_add(self.ctx, C.c_int64(first_id), C.c_int64(vectors.shape[0]), N.ptr(vectors)))

    def routing_build_add_dev(self, first_id: int, n: int, d_vectors: int):
        self._ck(self.lib.fspann_routing_build_add_dev(self.ctx, C.c_int64(first_id), C.c_int64(n), C.c_void_p(d_vectors)))

    def routing_build_finish(self, staged_ids=None, want_arrays: bool = False, n: int = 0):
        st = None if staged_ids is None else np.ascontiguousarray(staged_ids, dtype=np.int32)
        mn = mx = rep = ids = None
        if want_arrays:
            TD, P = self.T * self.D, (n + 63) // 64
            mn, mx = np.empty((TD, P), dtype=np.int64), np.empty((TD, P), dtype=np.int64)
            rep, ids = np.empty((TD, P, self.W), dtype=np.uint64), np.empty((TD, n), dtype=np.int32)
        self._ck(self.lib.fspann_routing_build_finish(self.ctx, N.ptr(st), N.ptr(mn), N.ptr(mx), N.ptr(rep), N.ptr(ids)))
        return mn, mx, rep, ids

    def store_alloc_shard(self, dim: int, id_base: int, n: int, n_global: int):
        self._ck(self.lib.fspann_store_alloc_shard(self.ctx, C.c_int64(id_base), C.c_int64(n), C.c_int64(n_global), C.c_int32(dim)))
        self.N, self.id_base, self.store_dim = n, id_base, dim

    def store_encrypt_dev(self, first_id: int, n: int, d_vectors: int, d_ivs: int, version: int):
        self._ck(self.lib.fspann_store_encrypt_dev(self.ctx, C.c_int64(first_id), C.c_int64(n), C.c_void_p(d_vectors), C.c_void_p(d_ivs), C.c_int32(version)))

    def deleted_set(self, flags):
        if flags is None:
            self._ck(self.lib.fspann_deleted_set(self.ctx, None, C.c_int64(0)))
        else:
            flags = np.ascontiguousarray(flags, dtype=np.uint8)
            self._ck(self.lib.fspann_deleted_set(self.ctx, N.ptr(flags), C.c_int64(flags.shape[0])))

    def store_upload(self, dim, iv, ct, key_version, id_base: int = 0, n_global: int | None = None):
        """Whole store (default) or the shard of global ids [id_base, id_base + len(iv)) of an n_global-record store."""
        iv = np.ascontiguousarray(iv, dtype=np.uint8)
        ct = np.ascontiguousarray(ct, dtype=np.uint8)
        key_version = np.ascontiguousarray(key_version, dtype=np.int32)
        n = iv.shape[0]
        assert iv.shape == (n, 12) and ct.shape == (n, 8 * dim + 16) and key_version.shape == (n,)
        if n_global is None:
            n_global = id_base + n
        self._ck(self.lib.fspann_store_upload_shard(self.ctx, C.c_int64(id_base), C.c_int64(n), C.c_int64(n_global), dim, N.ptr(iv), N.ptr(ct),
                                                    N.ptr(key_version)))
        self.N = n
        self.id_base = id_base
        self.store_dim = dim

    def store_update(self, ids, iv, ct, key_version):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        iv = np.ascontiguousarray(iv, dtype=np.uint8)
        ct = np.ascontiguousarray(ct, dtype=np.uint8)
        key_version = np.ascontiguousarray(key_version, dtype=np.int32)
        self._ck(self.lib.fspann_store_update(self.ctx, C.c_int64(ids.shape[0]), N.ptr(ids), N.ptr(iv), N.ptr(ct), N.ptr(key_version)))

    def migrate(self, ids, fresh_ivs, target_version: int):
        """reencryptTouched on the device (KRS:215-289).  Returns dict(reencrypted uint8[n], iv uint8[n,12], ct uint8[n,8d+16],
        count): rows of iv/ct are valid where reencrypted[i] == 1."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        fresh_ivs = np.ascontiguousarray(fresh_ivs, dtype=np.uint8)
        n = ids.shape[0]
        assert fresh_ivs.shape == (n, 12)
        done = np.zeros(n, dtype=np.uint8)
        iv = np.zeros((n, 12), dtype=np.uint8)
        ct = np.zeros((n, 8 * self.store_dim + 16), dtype=np.uint8)
        cnt = C.c_int64(0)
        self._ck(self.lib.fspann_migrate(self.ctx, C.c_int64(n), N.ptr(ids), N.ptr(fresh_ivs), C.c_int32(target_version), N.ptr(done),
                                         N.ptr(iv), N.ptr(ct), C.byref(cnt)))
        return dict(reencrypted=done, iv=iv, ct=ct, count=int(cnt.value))

    def encrypt_batch(self, ids, vectors, ivs, version: int):
        """encryptToPoint for a batch (AGC:55-112): FP64 [n, dim] -> ciphertext || tag uint8 [n, 8*dim+16]."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        vectors = np.ascontiguousarray(vectors, dtype=np.float64)
        ivs = np.ascontiguousarray(ivs, dtype=np.uint8)
        n, dim = vectors.shape
        assert ids.shape == (n,) and ivs.shape == (n, 12)
        ct = np.empty((n, 8 * dim + 16), dtype=np.uint8)
        self._ck(self.lib.fspann_encrypt_batch(self.ctx, C.c_int64(n), C.c_int32(dim), N.ptr(ids), N.ptr(vectors), N.ptr(ivs),
                                               C.c_int32(version), N.ptr(ct)))
        return ct

    def keys_set(self, version: int, key: bytes):
        assert len(key) == 32
        self._ck(self.lib.fspann_keys_set(self.ctx, C.c_int32(version), C.c_char_p(key)))

    def keys_retire(self, version: int):
        self._ck(self.lib.fspann_keys_retire(self.ctx, C.c_int32(version)))

    # ---- hot path, host buffers
    def tokengen_batch(self, queries):
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        if queries.ndim != 2 or queries.shape[1] != self.dim:
            raise N.IllegalArgumentError(f"Expected vector length {self.dim}")
        Q = queries.shape[0]
        codes = np.zeros((Q, self.T * self.D, self.W), dtype=np.uint64)
        self._ck(self.lib.fspann_tokengen_batch(self.ctx, C.c_int64(Q), N.ptr(queries), N.ptr(codes)))
        return codes

    def tokengen_batch_dev(self, Q, d_queries, d_codes):
        self._ck(self.lib.fspann_tokengen_batch_dev(self.ctx, C.c_int64(Q), C.c_void_p(d_queries), C.c_void_p(d_codes)))

    def route_batch(self, codes, probes, hard_cap, B, ham_threshold=0):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        Q = codes.shape[0]
        cid = np.full((Q, B), -1, dtype=np.int32)
        csc = np.full((Q, B), -1, dtype=np.int32)
        nc = np.zeros(Q, dtype=np.int32)
        raw = np.zeros(Q, dtype=np.int32)
        uq = np.zeros(Q, dtype=np.int32)
        self._ck(self.lib.fspann_route_batch(self.ctx, C.c_int64(Q), N.ptr(codes), C.c_int32(probes), C.c_int64(hard_cap),
                                             C.c_int32(ham_threshold), C.c_int32(B), N.ptr(cid), N.ptr(csc), N.ptr(nc), N.ptr(raw), N.ptr(uq)))
        return dict(cand_ids=cid, cand_scores=csc, n_cand=nc, raw_seen=raw, unique=uq)

    def refine_batch(self, queries, cand_ids, n_cand, k):
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        cand_ids = np.ascontiguousarray(cand_ids, dtype=np.int32)
        n_cand = np.ascontiguousarray(n_cand, dtype=np.int32)
        Q, stride = cand_ids.shape
        tid = np.full((Q, k), -1, dtype=np.int32)
        td = np.full((Q, k), np.nan, dtype=np.float64)
        nret = np.zeros(Q, dtype=np.int32)
        ver = np.full((Q, stride), 255, dtype=np.uint8)
        ndec = np.zeros(Q, dtype=np.int32)
        rank = np.full((Q, k), 0x7FFFFFFF, dtype=np.int32)
        self._ck(self.lib.fspann_refine_batch_ex(self.ctx, C.c_int64(Q), N.ptr(queries), N.ptr(cand_ids), N.ptr(n_cand), C.c_int32(stride),
                                                 C.c_int32(k), N.ptr(tid), N.ptr(td), N.ptr(rank), N.ptr(nret), N.ptr(ver), N.ptr(ndec)))
        return dict(top_ids=tid, top_dist=td, top_rank=rank, n_ret=nret, verdict=ver, n_decrypted=ndec)

    def search_batch(self, queries, k, probes, hard_cap, B, ham_threshold=0, out=None):
        """Host buffers in, host buffers out.  `queries` may be a numpy array or (address, Q) of pinned memory via out=."""
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        if queries.ndim != 2 or queries.shape[1] != self.dim:
            raise N.IllegalArgumentError(f"Query dimension mismatch: expected={self.dim} got={queries.shape[-1]}")
        Q = queries.shape[0]
        if out is None:
            out = dict(top_ids=np.full((Q, k), -1, dtype=np.int32), top_dist=np.full((Q, k), np.nan, dtype=np.float64),
                       n_ret=np.zeros(Q, dtype=np.int32), counters=np.zeros((Q, N.COUNTERS), dtype=np.int64))
        self._ck(self.lib.fspann_search_batch(self.ctx, C.c_int64(Q), N.ptr(queries), C.c_int32(k), C.c_int32(probes), C.c_int64(hard_cap),
                                              C.c_int32(B), C.c_int32(ham_threshold), N.ptr(out["top_ids"]), N.ptr(out["top_dist"]),
                                              N.ptr(out["n_ret"]), N.ptr(out["counters"])))
        return out

    def search_tokens(self, codes, queries, k, probes, hard_cap, B, ham_threshold=0):
        """QueryServiceImpl.search on the tokens' own codes (PIS:600): no TokenGen; a NaN/Inf query returns empty (QSI:137)."""
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        if queries.ndim != 2 or queries.shape[1] != self.dim:
            raise N.IllegalArgumentError(f"Query dimension mismatch: expected={self.dim} got={queries.shape[-1]}")
        Q = queries.shape[0]
        if codes is not None:
            codes = np.ascontiguousarray(codes, dtype=np.uint64).reshape(Q, self.T * self.D, self.W)
        out = dict(top_ids=np.full((Q, k), -1, dtype=np.int32), top_dist=np.full((Q, k), np.nan, dtype=np.float64),
                   n_ret=np.zeros(Q, dtype=np.int32), counters=np.zeros((Q, N.COUNTERS), dtype=np.int64))
        self._ck(self.lib.fspann_search_tokens(self.ctx, C.c_int64(Q), N.ptr(codes), N.ptr(queries), C.c_int32(k), C.c_int32(probes),
                                               C.c_int64(hard_cap), C.c_int32(B), C.c_int32(ham_threshold), N.ptr(out["top_ids"]),
                                               N.ptr(out["top_dist"]), N.ptr(out["n_ret"]), N.ptr(out["counters"])))
        return out

    def search_tokens_dev(self, Q, d_codes, d_queries, k, probes, hard_cap, B, ham_threshold, allow_retry, d_ids, d_dist, d_nret, d_counters=None):
        self._ck(self.lib.fspann_search_tokens_dev(self.ctx, C.c_int64(Q), C.c_void_p(d_codes), C.c_void_p(d_queries), C.c_int32(k),
                                                   C.c_int32(probes), C.c_int64(hard_cap), C.c_int32(B), C.c_int32(ham_threshold),
                                                   C.c_int32(allow_retry), C.c_void_p(d_ids), C.c_void_p(d_dist), C.c_void_p(d_nret),
                                                   C.c_void_p(d_counters) if d_counters else None))

    def search_batch_raw(self, Q, queries_addr, k, probes, hard_cap, B, ham_threshold, ids_addr, dist_addr, nret_addr, counters_addr):
        """Host-pointer call with caller-provided (e.g. pinned) buffers given as raw addresses."""
        self._ck(self.lib.fspann_search_batch(self.ctx, C.c_int64(Q), C.c_void_p(queries_addr), C.c_int32(k), C.c_int32(probes),
                                              C.c_int64(hard_cap), C.c_int32(B), C.c_int32(ham_threshold), C.c_void_p(ids_addr),
                                              C.c_void_p(dist_addr), C.c_void_p(nret_addr),
                                              C.c_void_p(counters_addr) if counters_addr else None))

    def search_batch_dev(self, Q, d_queries, k, probes, hard_cap, B, ham_threshold, allow_retry, d_ids, d_dist, d_nret, d_counters=None):
        """Device pointers (ints); enqueues on the context stream, no synchronisation unless allow_retry."""
        self._ck(self.lib.fspann_search_batch_dev(self.ctx, C.c_int64(Q), C.c_void_p(d_queries), C.c_int32(k), C.c_int32(probes),
                                                  C.c_int64(hard_cap), C.c_int32(B), C.c_int32(ham_threshold), C.c_int32(allow_retry),
                                                  C.c_void_p(d_ids), C.c_void_p(d_dist), C.c_void_p(d_nret),
                                                  C.c_void_p(d_counters) if d_counters else None))

    def groundtruth(self, base_f32, queries_f32, K: int, want_d2: bool = False):
        """GroundtruthPrecompute.run on the device: int32 [Q, K] ids ordered by (squared L2, id)."""
        base_f32 = np.ascontiguousarray(base_f32, dtype=np.float32)
        queries_f32 = np.ascontiguousarray(queries_f32, dtype=np.float32)
        n, dim = base_f32.shape
        Q = queries_f32.shape[0]
        if queries_f32.shape[1] != dim:
            raise N.IllegalArgumentError(f"Dim mismatch base={dim} query={queries_f32.shape[1]}")     # GroundtruthPrecompute:234-236
        K = min(max(1, K), n)                                                                     # kFinal (:238)
        ids = np.empty((Q, K), dtype=np.int32)
        d2 = np.empty((Q, K), dtype=np.float64) if want_d2 else None
        self._ck(self.lib.fspann_groundtruth(self.ctx, C.c_int64(n), C.c_int32(dim), N.ptr(base_f32), C.c_int64(Q), N.ptr(queries_f32),
                                             C.c_int32(K), N.ptr(ids), N.ptr(d2)))
        return (ids, d2) if want_d2 else ids

    def recall_batch(self, gt_ids, result_ids, K: int, n_ret=None) -> np.ndarray:
        """recall@K per query (FSA:785-794)."""
        gt_ids = np.ascontiguousarray(gt_ids, dtype=np.int32)
        result_ids = np.ascontiguousarray(result_ids, dtype=np.int32)
        Q = gt_ids.shape[0]
        nr = None if n_ret is None else np.ascontiguousarray(n_ret, dtype=np.int32)
        out = np.empty(Q, dtype=np.float64)
        self._ck(self.lib.fspann_recall_batch(self.ctx, C.c_int64(Q), C.c_int32(K), N.ptr(gt_ids), C.c_int32(gt_ids.shape[1]), N.ptr(result_ids),
                                              C.c_int32(result_ids.shape[1]), N.ptr(nr), N.ptr(out)))
        return out

    # ---- device-resident building blocks of the database-sharded deployment (all arguments are device addresses)
    def route_batch_dev(self, Q, d_queries, probes, hard_cap, B, d_cand_ids, d_n_cand, d_raw=None, d_unique=None):
        self._ck(self.lib.fspann_route_batch_dev(self.ctx, C.c_int64(Q), C.c_void_p(d_queries), C.c_int32(probes), C.c_int64(hard_cap), C.c_int32(B),
                                                 C.c_void_p(d_cand_ids), C.c_void_p(d_n_cand), C.c_void_p(d_raw) if d_raw else None,
                                                 C.c_void_p(d_unique) if d_unique else None))

    def refine_batch_dev(self, Q, d_queries, d_cand_ids, d_n_cand, stride, k, d_ids, d_dist, d_rank, d_nret, d_ndec=None):
        self._ck(self.lib.fspann_refine_batch_dev(self.ctx, C.c_int64(Q), C.c_void_p(d_queries), C.c_void_p(d_cand_ids), C.c_void_p(d_n_cand),
                                                  C.c_int32(stride), C.c_int32(k), C.c_void_p(d_ids), C.c_void_p(d_dist),
                                                  C.c_void_p(d_rank) if d_rank else None, C.c_void_p(d_nret), C.c_void_p(d_ndec) if d_ndec else None))

    def merge_topk_dev(self, S, Q, k, d_dist, d_rank, d_ids, d_out_ids, d_out_dist, d_out_nret):
        self._ck(self.lib.fspann_merge_topk_dev(self.ctx, C.c_int32(S), C.c_int64(Q), C.c_int32(k), C.c_void_p(d_dist), C.c_void_p(d_rank),
                                                C.c_void_p(d_ids), C.c_void_p(d_out_ids), C.c_void_p(d_out_dist), C.c_void_p(d_out_nret)))

    # ---- database-sharded search as one call: the NCCL collectives run inside the library (include/fspann_gpu.h)
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = self.lib.fspann_comm_unique_id(buf)
        if rc != N.OK:
            raise N.IllegalStateError("NCCL unavailable: fspann_comm_unique_id failed (libnccl.so.2 not loadable; set FSPANN_NCCL_LIB)")
        return bytes(buf)

    def comm_init(self, n_ranks: int, rank: int, comm_id: bytes | None = None):
        """COLLECTIVE over the n_ranks contexts holding the shards of one store (rank r holds the r-th id range)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(comm_id) if comm_id is not None else None
        self._ck(self.lib.fspann_comm_init(self.ctx, C.c_int32(n_ranks), C.c_int32(rank), buf))

    def comm_destroy(self):
        self._ck(self.lib.fspann_comm_destroy(self.ctx))

    def sharded_search_batch(self, queries, k, probes, hard_cap, B):
        """COLLECTIVE: every rank passes the same batch and gets the same (unsharded-identical) result."""
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        if queries.ndim != 2 or queries.shape[1] != self.dim:
            raise N.IllegalArgumentError(f"Query dimension mismatch: expected={self.dim} got={queries.shape[-1]}")
        Q = queries.shape[0]
        out = dict(top_ids=np.full((Q, k), -1, dtype=np.int32), top_dist=np.full((Q, k), np.nan, dtype=np.float64),
                   n_ret=np.zeros(Q, dtype=np.int32), counters=np.zeros((Q, N.COUNTERS), dtype=np.int64))
        self._ck(self.lib.fspann_sharded_search_batch(self.ctx, C.c_int64(Q), N.ptr(queries), C.c_int32(k), C.c_int32(probes), C.c_int64(hard_cap),
                                                      C.c_int32(B), N.ptr(out["top_ids"]), N.ptr(out["top_dist"]), N.ptr(out["n_ret"]),
                                                      N.ptr(out["counters"])))
        return out

    def sharded_search_batch_dev(self, Q, d_queries, k, probes, hard_cap, B, allow_retry, d_ids, d_dist, d_nret, d_counters=None):
        self._ck(self.lib.fspann_sharded_search_batch_dev(self.ctx, C.c_int64(Q), C.c_void_p(d_queries), C.c_int32(k), C.c_int32(probes),
                                                          C.c_int64(hard_cap), C.c_int32(B), C.c_int32(allow_retry), C.c_void_p(d_ids),
                                                          C.c_void_p(d_dist), C.c_void_p(d_nret), C.c_void_p(d_counters) if d_counters else None))

    def sharded_stage_ms(self):
        out = (C.c_float * 4)()
        nbytes = C.c_int64(0)
        n = self.lib.fspann_sharded_last_stage_ms(self.ctx, out, C.byref(nbytes))
        return dict(route=out[0], allgather_candidates=out[1], refine=out[2], allgather_topk_merge=out[3], gather_bytes=int(nbytes.value), launches=int(n))

    def touched(self, clear: bool = False) -> np.ndarray:
        words = (self.N + 31) // 32
        bm = np.zeros(words, dtype=np.uint32)
        self._ck(self.lib.fspann_touched_fetch(self.ctx, N.ptr(bm), C.c_int64(words), C.c_int32(1 if clear else 0)))
        bits = np.unpackbits(bm.view(np.uint8), bitorder="little")[: self.N]
        return (np.nonzero(bits)[0] + getattr(self, "id_base", 0)).astype(np.int32)

    def stage_ms(self):
        out = (C.c_float * 6)()
        n = self.lib.fspann_last_stage_ms(self.ctx, out)
        return dict(tokengen=out[0], route=out[1], group=out[2], verify=out[3], decrypt=out[4], topk=out[5], launches=int(n))

    def set_option(self, name: str, value: int):
        self._ck(self.lib.fspann_set_option(self.ctx, C.c_char_p(name.encode()), C.c_int64(value)))

    def get_info(self, name: str) -> int:
        return int(self.lib.fspann_get_info(self.ctx, C.c_char_p(name.encode())))

    def stream(self) -> int:
        return int(self.lib.fspann_ctx_stream(self.ctx) or 0)

    def sync(self):
        self._ck(self.lib.fspann_ctx_sync(self.ctx))

    def launch_count(self) -> int:
        return int(self.lib.fspann_ctx_launch_count(self.ctx))

    def debug_decrypt(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n = ids.shape[0]
        pt = np.zeros((n, self.store_dim), dtype=np.float64)
        ver = np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.fspann_debug_decrypt(self.ctx, C.c_int64(n), N.ptr(ids), N.ptr(pt), N.ptr(ver)))
        return pt, ver
