"""Host-side mirror of the reference's operator interface for the query hot path, backed by the CUDA library.

Same names, argument meaning and error behaviour as the Java classes, so the parity tests read like the
reference's own tests (no JDK in this image, so the host side above the C ABI is Python; the JNI / Panama
binding a Java maintainer would add is in INTEGRATION.md):

  SystemConfig                 config/src/main/java/com/fspann/config/SystemConfig.java:44-85,237-337
  QueryToken / QueryResult     common/src/main/java/com/fspann/common/QueryToken.java:23-71, QueryResult.java:6-23
  QueryTokenFactory.create     query/src/main/java/com/fspann/query/core/QueryTokenFactory.java:63-167
  PartitionedIndexService      index/src/main/java/com/fspann/index/paper/PartitionedIndexService.java (insert 266,
                               finalizeForSearch 789, lookupCandidatesWithScores 592, set/clearProbeOverride 868-874)
  QueryServiceImpl.search      query/src/main/java/com/fspann/query/service/QueryServiceImpl.java:100-352
  ForwardSecureANNSystem       api/src/main/java/com/fspann/api/ForwardSecureANNSystem.java (batchInsert 479,
                               finalizeForSearch 977, createToken 1673, runQueries 622)

TokenGen, Route and Refine run on the GPU through include/fspann_gpu.h; Setup / Rotate / Migrate / Retire stay on the
host (hostsetup.py), exactly as north_star prescribes.  Nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import hostsetup as HS
from ._native import IllegalArgumentError, IllegalStateError
from .gpu import GpuContext

DEFAULT_MAX_PROBES = 5  # PIS:93


@dataclass
class PaperConfig:          # CFG:237-263
    m: int = 24
    lam: int = 2
    divisions: int = 8
    tables: int = 8
    seed: int = 13


@dataclass
class RuntimeConfig:        # CFG:279-337
    refinementLimit: int = 1024
    maxGlobalCandidates: int = 20000
    probeOverride: int = -1
    hammingPrefilterThreshold: int = 0


@dataclass
class SystemConfig:
    paper: PaperConfig = field(default_factory=PaperConfig)
    runtime: RuntimeConfig = field(default_factory=RuntimeConfig)


@dataclass(frozen=True)
class QueryResult:          # common/.../QueryResult.java:6-23
    id: str
    distance: float


@dataclass
class QueryToken:           # QT:28-44
    bitCodes: np.ndarray    # uint64 [T, D, W]  (BitSet[table][division] as BitSet.toLongArray words)
    iv: bytes
    encryptedQuery: bytes
    topK: int
    numTables: int
    dimension: int
    version: int
    lam: int
    encryptionContext: str

    def getBitCodes(self):  # defensive copy like QT:76-80
        return self.bitCodes.copy()


class GFunctionRegistry:
    """Holder of the T*D GFunctions (GFR:40-147).  Values cross the boundary as data: they are built by
    the host (Java's SplittableRandom + Math.log/cos) and uploaded; this class only stores and validates them."""

    def __init__(self):
        self.initialized = False

    def initialize(self, dim, m, lam, tables, divisions, alpha, r, omega):
        alpha = np.ascontiguousarray(alpha, dtype=np.float64).reshape(tables * divisions, m, dim)
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(tables * divisions, m)
        omega = np.ascontiguousarray(omega, dtype=np.float64).reshape(tables * divisions, m)
        if not np.all(omega > 0):
            raise IllegalArgumentError("omega_j <= 0")        # Coding:85-87
        self.dim, self.m, self.lam, self.tables, self.divisions = dim, m, lam, tables, divisions
        self.alpha, self.r, self.omega = alpha, r, omega
        self.initialized = True

    def reset(self):
        self.initialized = False

    def stats(self):
        return dict(dimension=self.dim, m=self.m, lam=self.lam, tables=self.tables, divisions=self.divisions)


class PartitionedIndexService:
    """GPU-backed PartitionedIndexService: Setup keeps the reference's staging semantics on the host, Route runs in
    route.cu.  `insert` takes the already-encrypted record (the Java host encrypts with its own KeyManager)."""

    def __init__(self, gpu: GpuContext, cfg: SystemConfig, registry: GFunctionRegistry):
        self.gpu, self.cfg, self.registry = gpu, cfg, registry
        self.frozen = False
        self._ids: list[int] = []
        self._vecs: list[np.ndarray] = []
        self._probe_override = -1
        self.lastRawVisited = 0
        self.lastTouchedIds: list[int] = []

    # -- Setup
    def insert(self, id_: int, vector: np.ndarray):
        if self.frozen:
            raise IllegalStateError("Index already finalized")
        if id_ is None or vector is None:
            raise IllegalArgumentError("id / vector cannot be null")
        v = np.asarray(vector, dtype=np.float64)
        if self._vecs and v.shape[0] != self._vecs[0].shape[0]:
            raise IllegalArgumentError(f"Mixed dimensions not supported in single index: got {v.shape[0]}, expected {self._vecs[0].shape[0]}")
        self._ids.append(int(id_))
        self._vecs.append(v)

    def insert_many(self, ids: np.ndarray, vectors: np.ndarray):
        """Bulk form of insert(): APPENDS to what is already staged (a second batchInsert continues the index, FSA:501,515)."""
        if self.frozen:
            raise IllegalStateError("Index already finalized")
        ids, vectors = np.asarray(ids, dtype=np.int32), np.ascontiguousarray(vectors, dtype=np.float64)
        if hasattr(self, "_bulk"):
            if vectors.shape[1] != self._bulk[1].shape[1]:
                raise IllegalArgumentError(f"Mixed dimensions not supported in single index: got {vectors.shape[1]}, expected {self._bulk[1].shape[1]}")
            ids, vectors = np.concatenate([self._bulk[0], ids]), np.concatenate([self._bulk[1], vectors])
        self._bulk = (ids, vectors)

    def finalizeForSearch(self):
        """PIS:789-845: code every staged vector for every (t,d) (on the GPU, same kernel as TokenGen), build the greedy
        partitions in staged order, freeze, upload the routing state."""
        if self.frozen:
            return
        ids, vecs = np.asarray(self._ids, dtype=np.int32), (np.stack(self._vecs) if self._vecs else None)
        if hasattr(self, "_bulk"):                       # bulk-staged rows come first, single insert()s after them
            ids = np.concatenate([self._bulk[0], ids])
            vecs = self._bulk[1] if vecs is None else np.concatenate([self._bulk[1], vecs])
        if vecs is None:
            vecs = np.zeros((0, 1))
        n = ids.shape[0]
        if n < HS.MIN_SAMPLE_SIZE:
            raise IllegalStateError(f"Cannot finalize index: only {n} samples collected (< MIN_SAMPLE_SIZE)")   # PIS:803-808
        reg, pc = self.registry, self.cfg.paper
        if not reg.initialized:
            raise IllegalStateError("GFunctionRegistry not initialized")
        st = reg.stats()
        if (st["m"], st["lam"], st["tables"], st["divisions"]) != (pc.m, pc.lam, pc.tables, pc.divisions):
            raise IllegalStateError(f"GFunctionRegistry mismatch at finalize: {st}")                             # PIS:812-819
        if vecs.shape[1] != reg.dim:
            raise IllegalArgumentError(f"Mixed dimensions not supported in single index: got {vecs.shape[1]}, expected {reg.dim}")
        if not np.array_equal(np.sort(ids), np.arange(n, dtype=np.int32)):
            raise IllegalArgumentError("ids must be the ordinals 0..N-1 (FSA:501,515)")
        self.gpu.gfunctions_upload(reg.dim, pc.tables, pc.divisions, pc.m, pc.lam, reg.alpha, reg.r, reg.omega)
        inv = np.empty(n, dtype=np.int64)
        inv[ids] = np.arange(n)
        # staged order: in insertion order, the 1000th and later first, then the 999 parked ones (PIS:280-298, 821-831)
        staged_ids = ids[HS.staged_order(n)]
        # coding loop + GreedyPartitioner.build for every (t,d) on the device; the result is installed as the routing state
        mn, mx, rep, pids = self.gpu.routing_build(vecs[inv], staged_ids)
        self.routing = HS.RoutingIndex(reg.dim, pc.tables, pc.divisions, pc.m, pc.lam, reg.alpha, reg.r, reg.omega, n, mn, mx, rep, pids)
        self.frozen = True
        self._ids, self._vecs = [], []

    def isFrozen(self):
        return self.frozen

    def numTables(self):
        return self.cfg.paper.tables

    # -- probes (PIS:868-888)
    def setProbeOverride(self, probes: int):
        self._probe_override = probes

    def clearProbeOverride(self):
        self._probe_override = -1

    def getDefaultMaxProbes(self):
        return DEFAULT_MAX_PROBES

    def effectiveMaxProbes(self):
        if self._probe_override > 0:
            return self._probe_override
        if self.cfg.runtime.probeOverride > 0:
            return self.cfg.runtime.probeOverride
        return DEFAULT_MAX_PROBES

    def hardCap(self):          # PIS:612-615
        return max(self.cfg.runtime.maxGlobalCandidates, self.cfg.runtime.refinementLimit)

    # -- Route
    def lookupCandidatesWithScores(self, token: QueryToken, limit: int | None = None):
        """PIS:592-715.  Returns [(id, hammingDist)] of every unique candidate in the reference's order (or the first `limit`);
        getLastRawCandidateCount / getLastTouchedIds / getLastTouchedCount cover all of them, as in the reference."""
        if token is None:
            raise IllegalArgumentError("token")
        if not self.frozen:
            raise IllegalStateError("Index not finalized")
        if token.bitCodes is None:
            raise IllegalStateError("MSANNP violation: QueryToken missing BitSet codes")
        if token.bitCodes.shape[0] != self.cfg.paper.tables:
            raise IllegalStateError(f"Token tables mismatch: token={token.bitCodes.shape[0]} index={self.cfg.paper.tables}")
        if token.dimension != self.gpu.dim:
            return []
        codes = token.bitCodes.reshape(1, -1, self.gpu.W)
        # the reference returns ALL unique candidates (<= HARD_CAP + 63), score-sorted, and records every one of them as touched
        # (PIS:690-702: result.size() = bestScore.size()); `limit` only trims what this call hands back
        full = min(self.hardCap() + 64, self.cfg.paper.tables * self.cfg.paper.divisions * max(self.effectiveMaxProbes(), 1) * 64)
        out = self.gpu.route_batch(codes, self.effectiveMaxProbes(), self.hardCap(), max(full, 1), self.cfg.runtime.hammingPrefilterThreshold)
        n = int(out["n_cand"][0])
        self.lastRawVisited = int(out["raw_seen"][0])
        self.lastUnique = int(out["unique"][0])
        self.lastTouchedIds = out["cand_ids"][0, :n].tolist()
        if limit is not None:
            n = min(n, limit)
        return list(zip(out["cand_ids"][0, :n].tolist(), out["cand_scores"][0, :n].tolist()))

    def lookupCandidateIds(self, token: QueryToken, limit: int | None = None):
        """PIS:459-582, the ids-only twin of lookupCandidatesWithScores (same traversal and order)."""
        return [i for i, _ in self.lookupCandidatesWithScores(token, limit)]

    def getLastRawCandidateCount(self):
        return self.lastRawVisited

    def getLastTouchedIds(self):        # PIS:860-866
        return list(self.lastTouchedIds)

    def getLastTouchedCount(self):
        return len(self.lastTouchedIds)


class QueryTokenFactory:
    def __init__(self, gpu: GpuContext, keys: HS.KeyManager, cfg: SystemConfig, registry: GFunctionRegistry, rng=None):
        """rng: TEST-ONLY deterministic IV source (an object with .bytes(n)).  The default draws every 96-bit IV from the operating
        system's CSPRNG like the reference's SecureRandom (QTF:152-154): a (key, IV) pair must never repeat under AES-GCM."""
        self.gpu, self.keys, self.cfg, self.registry = gpu, keys, cfg, registry
        self.rng = rng

    def create(self, vec, topK: int) -> QueryToken:
        if vec is None:
            raise IllegalArgumentError("query vector is null")
        if topK <= 0:
            raise IllegalArgumentError("topK must be > 0")
        if not self.registry.initialized:
            raise IllegalStateError("GFunctionRegistry not initialized. Build index first.")
        vec = np.asarray(vec, dtype=np.float64)
        pc, st = self.cfg.paper, self.registry.stats()
        if (st["dimension"], st["tables"], st["divisions"], st["m"], st["lam"]) != (vec.shape[0], pc.tables, pc.divisions, pc.m, pc.lam):
            raise IllegalStateError(f"GFunctionRegistry mismatch: {st}")
        codes = self.gpu.tokengen_batch(vec[None, :])[0].reshape(pc.tables, pc.divisions, self.gpu.W)
        version = self.keys.current
        iv = self.rng.bytes(12) if self.rng is not None else os.urandom(12)
        ct = HS.encrypt_query(vec, self.keys.get_version(version), iv)
        return QueryToken(codes, iv, ct, max(1, topK), max(1, pc.tables), vec.shape[0], version, pc.lam, f"dim_{vec.shape[0]}_v{version}")

    def derive(self, tok: QueryToken, newTopK: int) -> QueryToken:
        """QTF:182-200: the same token (codes, IV, ciphertext, version) with another topK -- no new TokenGen, no new encryption."""
        if tok is None:
            raise IllegalArgumentError("token is null")
        if newTopK <= 0:
            raise IllegalArgumentError("newTopK must be > 0")
        return QueryToken(tok.getBitCodes(), tok.iv, tok.encryptedQuery, newTopK, tok.numTables, tok.dimension, tok.version, tok.lam,
                          tok.encryptionContext)


class QueryServiceImpl:
    """QueryService.search on the GPU.  `search(token)` keeps the reference's one-token signature; `searchBatch` is
    the batched form a GPU host should use."""

    def __init__(self, index: PartitionedIndexService, gpu: GpuContext, keys: HS.KeyManager, cfg: SystemConfig):
        self.index, self.gpu, self.keys, self.cfg = index, gpu, keys, cfg
        self._refine_override = None
        self.lastCandTotal = self.lastCandKept = self.lastCandDecrypted = self.lastReturned = 0
        self.lastRetried = False
        self.touchedThisSession: set[int] = set()

    def setRefinementLimit(self, limit: int):
        self._refine_override = limit

    def clearRefinementLimit(self):
        self._refine_override = None

    def getEffectiveRefinementLimit(self):
        return self._refine_override if (self._refine_override or 0) > 0 else self.cfg.runtime.refinementLimit

    def _decrypt_token(self, token: QueryToken) -> np.ndarray | None:
        try:
            key = self.keys.get_version(token.version)          # QSI:124-129
        except Exception:
            key = self.keys.get_version(self.keys.current)
        q = HS.decrypt_query(token.encryptedQuery, key, token.iv)
        if not np.all(np.isfinite(q)):                          # QSI:137
            return None
        return q

    def search(self, token: QueryToken):
        if token is None:
            return []
        res = self.searchBatch([token])
        return res[0]

    def searchBatch(self, tokens):
        """QueryServiceImpl.search for a list of tokens.  Routes on each token's OWN codes (token.getBitCodes(), PIS:600) through
        fspann_search_tokens -- TokenGen ran once, when the token was created -- and refines against the decrypted query."""
        if not self.index.isFrozen():
            raise IllegalStateError("Index not finalized")
        self.touchedThisSession = set()
        qs, cs, ks, keep = [], [], [], []
        for i, t in enumerate(tokens):
            if t is None:
                continue
            if t.bitCodes is None:
                raise IllegalStateError("MSANNP violation: QueryToken missing BitSet codes")       # PIS:604-606
            if t.bitCodes.shape[0] != self.cfg.paper.tables:
                raise IllegalStateError(f"Token tables mismatch: token={t.bitCodes.shape[0]} index={self.cfg.paper.tables}")
            q = self._decrypt_token(t)
            if q is not None and t.dimension == self.gpu.dim:
                qs.append(q); cs.append(t.bitCodes.reshape(-1, self.gpu.W)); ks.append(t.topK); keep.append(i)
        results = [[] for _ in tokens]
        if not qs:
            return results
        k = max(ks)
        if len(set(ks)) != 1:
            raise IllegalArgumentError("searchBatch needs one topK per batch (derive tokens per K like FSA:634)")
        rt = self.cfg.runtime
        try:
            out = self.gpu.search_tokens(np.stack(cs), np.stack(qs), k, self.index.effectiveMaxProbes(), self.index.hardCap(),
                                         self.getEffectiveRefinementLimit(), rt.hammingPrefilterThreshold)
        finally:
            self.index.clearProbeOverride()                       # QSI:342-346: cleared in `finally`, also when the search throws
        for j, i in enumerate(keep):
            n = int(out["n_ret"][j])
            results[i] = [QueryResult(str(int(out["top_ids"][j, r])), float(out["top_dist"][j, r])) for r in range(n)]
        c = out["counters"][-1]
        self.lastCandTotal, self.lastCandKept, self.lastCandDecrypted, self.lastReturned = int(c[0]), int(c[1]), int(c[2]), int(c[3])
        self.lastRetried = bool(c[4])
        self.lastCounters = out["counters"]
        self.touchedThisSession = set(self.gpu.touched(clear=True).tolist())   # reencTracker.record(touched) (QSI:348-350)
        return results

    def getLastCandTotal(self): return self.lastCandTotal
    def getLastCandKept(self): return self.lastCandKept
    def getLastCandDecrypted(self): return self.lastCandDecrypted
    def getLastReturned(self): return self.lastReturned


class ForwardSecureANNSystem:
    """Facade for the path only: batchInsert -> finalizeForSearch -> createToken -> search (FSA:479, 977, 1673, 622).
    Keys are derived by the host KeyManager; bulk encryption at insert, the index build and Migrate run on the device
     and the host keeps the persistent mirror of the encrypted store."""

    def __init__(self, cfg: SystemConfig, dim: int, master_key: bytes, gfunctions, device: int = 0, iv_seed: int | None = None,
                 debug: bool = False):
        """iv_seed=None (default): every record and query IV comes from the OS CSPRNG, like the reference's SecureRandom (AGC:66-67,
        QTF:152-154).  iv_seed=<int> is a TEST-ONLY switch that makes the IVs reproducible; never use it with a real master key --
        two runs would then encrypt different data under the same (key, IV), which breaks AES-GCM."""
        self.cfg, self.dim = cfg, dim
        self.gpu = GpuContext(device, debug=debug)
        self.keys = HS.KeyManager(master_key)
        self.registry = GFunctionRegistry()
        pc = cfg.paper
        alpha, r, omega = gfunctions
        self.registry.initialize(dim, pc.m, pc.lam, pc.tables, pc.divisions, alpha, r, omega)
        self.index = PartitionedIndexService(self.gpu, cfg, self.registry)
        self.tokenFactory = QueryTokenFactory(self.gpu, self.keys, cfg, self.registry,
                                              None if iv_seed is None else np.random.default_rng(iv_seed + 1))
        self.queryService = QueryServiceImpl(self.index, self.gpu, self.keys, cfg)
        self.iv_seed = iv_seed
        self.store_iv = self.store_ct = self.store_ver = None
        self.gpu.keys_set(1, self.keys.derive(1))

    def batchInsert(self, vectors: np.ndarray, ivs: np.ndarray | None = None):
        """FSA:479-560: ids continue the ordinals of what was inserted before (FSA:501,515), so a second call APPENDS.  `ivs`
        (uint8 [n,12]) must come from a CSPRNG; when omitted they are drawn from os.urandom (or the test-only seeded generator)."""
        vectors = np.ascontiguousarray(vectors, dtype=np.float64)
        if vectors.shape[1] != self.dim:
            raise IllegalArgumentError(f"Expected vector length {self.dim}")
        if self.index.isFrozen():
            raise IllegalStateError("Index already finalized")
        n = vectors.shape[0]
        n0 = 0 if getattr(self, "store_iv", None) is None else self.store_iv.shape[0]
        ids = np.arange(n0, n0 + n, dtype=np.int32)                # FSA:501,515 id = ordinal
        if ivs is None:
            if self.iv_seed is None:
                ivs = np.frombuffer(os.urandom(12 * n), dtype=np.uint8).reshape(n, 12)
            else:
                from . import workloads as WL
                ivs = WL.record_ivs(n, self.iv_seed + n0)
        v = self.keys.current
        ivs = np.ascontiguousarray(ivs, dtype=np.uint8).copy()
        ct = self.gpu.encrypt_batch(ids, vectors, ivs, v)          # encryptToPoint (AGC:55-112) on the device
        if n0:
            self.store_iv, self.store_ct = np.concatenate([self.store_iv, ivs]), np.concatenate([self.store_ct, ct])
            self.store_ver = np.concatenate([self.store_ver, np.full(n, v, dtype=np.int32)])
        else:
            self.store_iv, self.store_ct, self.store_ver = ivs, ct, np.full(n, v, dtype=np.int32)
        self.gpu.store_upload(self.dim, self.store_iv, self.store_ct, self.store_ver)
        self.index.insert_many(ids, vectors)

    def finalizeForSearch(self):
        self.index.finalizeForSearch()

    def createToken(self, q, topK: int, dim: int) -> QueryToken:   # FSA:1673-1698
        if not self.index.isFrozen():
            raise IllegalStateError("Index is not finalized; call finalizeForSearch() before querying")
        if q is None or len(q) != dim or dim != self.dim:
            raise IllegalArgumentError(f"Query dimension mismatch: expected={self.dim} got={None if q is None else len(q)}")
        return self.tokenFactory.create(q, topK)

    # -- callers of the path
    def evalSimple(self, q, topK: int, dim: int):                   # QueryFacade.evalSimple (FSA:863-875)
        return self.queryService.search(self.createToken(q, topK, dim))

    def runQueries(self, queries, dim: int, gt_ids=None, k_variants=(1, 10, 20, 40, 60, 80, 100)):
        """FSA:622-748 for a batch: one token per query at MAX_K = max(kVariants) (FSA:634,661), search, the probes-only fallback
        max(2p, 4) for queries that came back empty (FSA:667-678), then per-K prefix metrics.  gt_ids: int32 [Q, >= MAX_K] ground truth
        (fspann_groundtruth / GroundtruthPrecompute) or None.  Returns dict(results=[[QueryResult]], recall={K: mean recall@K},
        returned={K: mean |prefix|}, fallback=[query indices that needed the fallback])."""
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        max_k = int(max(k_variants))
        rt = self.cfg.runtime
        if rt.probeOverride >= 0:
            self.index.setProbeOverride(rt.probeOverride)            # FSA:641-644
        if rt.refinementLimit > 0:
            self.queryService.setRefinementLimit(rt.refinementLimit)
        tokens = [self.createToken(queries[i], max_k, dim) for i in range(queries.shape[0])]
        results = self.queryService.searchBatch(tokens)
        empty = [i for i, r in enumerate(results) if not r]
        if empty:
            base = rt.probeOverride if rt.probeOverride >= 0 else self.index.getDefaultMaxProbes()
            self.index.setProbeOverride(max(base * 2, 4))
            again = self.queryService.searchBatch([tokens[i] for i in empty])
            self.index.clearProbeOverride()                          # "DO NOT touch refinement limit" (FSA:677)
            for i, r in zip(empty, again):
                results[i] = r
        self.index.clearProbeOverride()
        self.queryService.clearRefinementLimit()
        recall, returned = {}, {}
        Q = queries.shape[0]
        ids = np.full((Q, max_k), -1, dtype=np.int32)
        nret = np.zeros(Q, dtype=np.int32)
        for i, r in enumerate(results):
            nret[i] = len(r)
            ids[i, :len(r)] = [int(x.id) for x in r]
        for k in k_variants:
            returned[k] = float(np.minimum(nret, k).mean())
            if gt_ids is not None and gt_ids.shape[1] >= k:
                recall[k] = float(self.gpu.recall_batch(np.ascontiguousarray(gt_ids[:, :k]), ids, k, np.minimum(nret, k)).mean())   # FSA:785-794
        return dict(results=results, recall=recall, returned=returned, fallback=empty)

    # -- lifecycle driven from the host (Rotate / Migrate / Retire), mirrored into the GPU state
    def rotateKeyOnly(self) -> int:
        v = self.keys.rotate_key_only()
        self.gpu.keys_set(v, self.keys.derive(v))
        return v

    def reencryptTouched(self, ids, fresh_ivs, target_version: int, on_device: bool = True):
        """KRS:215-289.  on_device=True: Migrate runs on the GPU in place in the HBM store (fspann_migrate) and the new records
        come back for the host's persistent copy; on_device=False: the host re-encrypts (the reference's own arrangement) and
        mirrors the result with fspann_store_update."""
        ids = np.asarray(ids, dtype=np.int32)
        if not on_device:
            done = HS.migrate(self.store_iv, self.store_ct, self.store_ver, ids, fresh_ivs, target_version, self.keys)
            if done:
                d = np.asarray(done, dtype=np.int32)
                self.gpu.store_update(d, self.store_iv[d], self.store_ct[d], self.store_ver[d])
            return done
        out = self.gpu.migrate(ids, fresh_ivs, target_version)
        sel = np.nonzero(out["reencrypted"])[0]
        d = ids[sel]
        self.store_iv[d], self.store_ct[d], self.store_ver[d] = out["iv"][sel], out["ct"][sel], target_version   # saveEncryptedPoint (KRS:268)
        return d.tolist()

    def retire(self, version: int) -> bool:
        if np.any(self.store_ver == version):                       # KM:287-294: refuse while vectors are still bound
            return False
        self.keys.retire(version)
        self.gpu.keys_retire(version)
        return True

    def shutdown(self):
        self.gpu.close()
