"""ctypes binding of oracle/libfspann_oracle.so -- the CPU restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY (see the header of fspann_oracle.c).  Importable from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs, never from the
product package.  Parity status: unpinned by reference fixtures (none exist); AES-GCM and
HMAC are pinned against NIST / RFC known-answer vectors in tests/test_oracle_crypto.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfspann_oracle.so")

BLOCK = 64
VERDICT_OK, VERDICT_NOT_FOUND, VERDICT_NO_KEY, VERDICT_TAG_FAIL, VERDICT_NON_FINITE = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fspann_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfspann_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_compute_key.restype = C.c_int64
        _lib.orc_hamming.restype = C.c_int64
        _lib.orc_find_nearest.restype = C.c_int64
        _lib.orc_route.restype = C.c_int64
        _lib.orc_migrate.restype = C.c_int64
        _lib.orc_java_hash_decimal.restype = C.c_uint32
        _lib.orc_hashmap_final_cap.restype = C.c_uint32
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class _IndexStruct(C.Structure):
    _fields_ = [("dim", C.c_int32), ("T", C.c_int32), ("D", C.c_int32), ("m", C.c_int32), ("lam", C.c_int32),
                ("W", C.c_int32), ("N", C.c_int64), ("P", C.c_int64), ("min_key", C.c_void_p), ("max_key", C.c_void_p),
                ("rep", C.c_void_p), ("ids", C.c_void_p), ("deleted", C.c_void_p), ("n_deleted_flags", C.c_int64)]


class _StoreStruct(C.Structure):
    _fields_ = [("N", C.c_int64), ("dim", C.c_int32), ("iv", C.c_void_p), ("ct", C.c_void_p), ("key_version", C.c_void_p),
                ("deleted", C.c_void_p), ("present", C.c_void_p), ("n_keys", C.c_int32), ("key_versions", C.c_void_p),
                ("keys", C.c_void_p)]


@dataclass
class GFunctions:
    """All T*D GFunctions as flat FP64 arrays (Coding.GFunction fields, Coding:52-97)."""
    dim: int
    T: int
    D: int
    m: int
    lam: int
    alpha: np.ndarray  # [T*D, m, dim]
    r: np.ndarray      # [T*D, m]
    omega: np.ndarray  # [T*D, m]

    @property
    def W(self) -> int:
        return (self.m * self.lam + 63) // 64


def registry_init(sample: np.ndarray, m: int, lam: int, seed: int, T: int, D: int) -> GFunctions:
    """GFunctionRegistry.initialize (GFR:63-147)."""
    sample = np.ascontiguousarray(sample, dtype=np.float64)
    n, d = sample.shape
    alpha = np.empty((T * D, m, d), dtype=np.float64)
    r = np.empty((T * D, m), dtype=np.float64)
    omega = np.empty((T * D, m), dtype=np.float64)
    lib().orc_registry_init(_p(sample), C.c_int(n), C.c_int(d), C.c_int(m), C.c_int(lam), C.c_int64(seed), C.c_int(T),
                            C.c_int(D), _p(alpha), _p(r), _p(omega))
    return GFunctions(d, T, D, m, lam, alpha, r, omega)


def H(v: np.ndarray, g: GFunctions, td: int) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty(g.m, dtype=np.int32)
    lib().orc_H(_p(v), C.c_int(g.dim), C.c_int(g.m), _p(g.alpha[td]), _p(g.r[td]), _p(g.omega[td]), _p(out))
    return out


def tokengen_batch(queries: np.ndarray, g: GFunctions) -> np.ndarray:
    """Coding.C for every (t,d) of every query (QTF:98-131).  Returns uint64 [Q, T*D, W]."""
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    Q = queries.shape[0]
    codes = np.zeros((Q, g.T * g.D, g.W), dtype=np.uint64)
    lib().orc_tokengen_batch(_p(queries), C.c_int64(Q), C.c_int(g.dim), C.c_int(g.m), C.c_int(g.lam), C.c_int(g.T),
                             C.c_int(g.D), _p(g.alpha), _p(g.r), _p(g.omega), _p(codes), C.c_int(g.W))
    return codes


def compute_key(code: np.ndarray) -> int:
    code = np.ascontiguousarray(code, dtype=np.uint64)
    return int(lib().orc_compute_key(_p(code), C.c_int(code.shape[0])))


def java_hash_decimal(i: int) -> int:
    return int(lib().orc_java_hash_decimal(C.c_int64(i)))


def hashmap_order(keys: np.ndarray, initial_capacity: int):
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    order = np.empty(keys.shape[0], dtype=np.int32)
    mc = lib().orc_hashmap_order(_p(keys), C.c_int64(keys.shape[0]), C.c_int64(initial_capacity), _p(order))
    return order, int(mc)


@dataclass
class Index:
    """Flat routing index: the arrays fspann_routing_upload takes (include/fspann_gpu.h)."""
    g: GFunctions
    N: int
    P: int
    min_key: np.ndarray  # int64 [T*D, P]
    max_key: np.ndarray  # int64 [T*D, P]
    rep: np.ndarray      # uint64 [T*D, P, W]
    ids: np.ndarray      # int32 [T*D, N]
    deleted: np.ndarray | None = None  # uint8 [>= max id + 1]
    max_chain: int = 0
    _keep: list = field(default_factory=list)

    def struct(self) -> _IndexStruct:
        g = self.g
        s = _IndexStruct(g.dim, g.T, g.D, g.m, g.lam, g.W, self.N, self.P, self.min_key.ctypes.data, self.max_key.ctypes.data,
                         self.rep.ctypes.data, self.ids.ctypes.data,
                         self.deleted.ctypes.data if self.deleted is not None else None,
                         self.deleted.shape[0] if self.deleted is not None else 0)
        return s


def staged_order(N: int) -> np.ndarray:
    """Order in which PIS stages vectors: ids 999..N-1 as inserted, then the 999 parked ones (PIS:280-298,821-831)."""
    if N < 1000:
        raise ValueError("reference refuses to initialise with < 1000 vectors (PIS:50,171-177)")
    return np.concatenate([np.arange(999, N, dtype=np.int32), np.arange(0, 999, dtype=np.int32)])


def index_build(codes_by_id: np.ndarray, g: GFunctions, staged_ids: np.ndarray) -> Index:
    """GreedyPartitioner.build for every (t,d) (PIS:412-425 -> GP:37-76).
    codes_by_id: uint64 [N, T*D, W] indexed by id; staged_ids: insertion order into the HashMap."""
    N = staged_ids.shape[0]
    TD, W = g.T * g.D, g.W
    P = (N + BLOCK - 1) // BLOCK
    staged_ids = np.ascontiguousarray(staged_ids, dtype=np.int32)
    staged_codes = np.ascontiguousarray(codes_by_id[staged_ids])  # [N, TD, W] in staged order
    mn = np.zeros((TD, P), dtype=np.int64)
    mx = np.zeros((TD, P), dtype=np.int64)
    rep = np.zeros((TD, P, W), dtype=np.uint64)
    ids = np.zeros((TD, N), dtype=np.int32)
    max_chain = 0
    for td in range(TD):
        base = staged_codes.ctypes.data + td * W * 8
        mc = lib().orc_partition_build(_p(staged_ids), C.c_void_p(base), C.c_int64(N), C.c_int(W), C.c_int64(TD * W),
                                       _p(mn[td]), _p(mx[td]), _p(rep[td]), _p(ids[td]))
        max_chain = max(max_chain, int(mc))
    return Index(g, N, P, mn, mx, rep, ids, None, max_chain)


def route(ix: Index, codes: np.ndarray, probes: int, hard_cap: int, max_out: int | None = None):
    """lookupCandidatesWithScores (PIS:592-715).  Returns (ids, scores, raw_seen, max_chain), all unique candidates."""
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    if max_out is None:
        max_out = ix.g.T * ix.g.D * max(probes, 0) * BLOCK + hard_cap + 2 * BLOCK
    oid = np.empty(max_out, dtype=np.int32)
    osc = np.empty(max_out, dtype=np.int32)
    raw = C.c_int64(0)
    mc = C.c_int(0)
    st = ix.struct()
    n = lib().orc_route(C.byref(st), _p(codes), C.c_int(probes), C.c_int64(hard_cap), _p(oid), _p(osc), C.c_int64(max_out),
                        C.byref(raw), C.byref(mc))
    n = min(int(n), max_out)
    return oid[:n].copy(), osc[:n].copy(), int(raw.value), int(mc.value)


# ---------------------------------------------------------------- crypto
def kdf(master: bytes, version: int) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_kdf(C.c_char_p(master), C.c_int32(version), out)
    return out.raw


def hmac_sha256(key: bytes, msg: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_hmac_sha256(C.c_char_p(key), C.c_int(len(key)), C.c_char_p(msg), C.c_int(len(msg)), out)
    return out.raw


def gcm_encrypt(key: bytes, iv: bytes, aad: bytes, pt: bytes) -> bytes:
    out = C.create_string_buffer(len(pt) + 16)
    rc = lib().orc_gcm_encrypt(C.c_char_p(key), C.c_char_p(iv), C.c_char_p(aad), C.c_int(len(aad)), C.c_char_p(pt),
                               C.c_int(len(pt)), out)
    assert rc == 0
    return out.raw


def gcm_decrypt(key: bytes, iv: bytes, aad: bytes, ct: bytes):
    out = C.create_string_buffer(max(len(ct), 1) + 16)
    rc = lib().orc_gcm_decrypt(C.c_char_p(key), C.c_char_p(iv), C.c_char_p(aad), C.c_int(len(aad)), C.c_char_p(ct),
                               C.c_int(len(ct)), out)
    return rc, out.raw[:max(len(ct) - 16, 0)]


def aad(id_: int, key_version: int, dim: int) -> bytes:
    buf = C.create_string_buffer(64)
    n = lib().orc_aad(C.c_int64(id_), C.c_int32(key_version), C.c_int32(dim), buf)
    return buf.raw[:n]


def encrypt_point(id_: int, key_version: int, vec: np.ndarray, key: bytes, iv: bytes) -> bytes:
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    out = C.create_string_buffer(vec.shape[0] * 8 + 16)
    rc = lib().orc_encrypt_point(C.c_int64(id_), C.c_int32(key_version), _p(vec), C.c_int(vec.shape[0]), C.c_char_p(key),
                                 C.c_char_p(iv), out)
    assert rc == 0
    return out.raw


def decrypt_point(id_: int, key_version: int, dim: int, key: bytes, iv: bytes, ct: bytes):
    out = np.empty(dim, dtype=np.float64)
    rc = lib().orc_decrypt_point(C.c_int64(id_), C.c_int32(key_version), C.c_int(dim), C.c_char_p(key), C.c_char_p(iv),
                                 C.c_char_p(ct), _p(out))
    return rc, out


def encrypt_query(vec: np.ndarray, key: bytes, iv: bytes) -> bytes:
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    out = C.create_string_buffer(vec.shape[0] * 8 + 16)
    assert lib().orc_encrypt_query(_p(vec), C.c_int(vec.shape[0]), C.c_char_p(key), C.c_char_p(iv), out) == 0
    return out.raw


def decrypt_query(ct: bytes, key: bytes, iv: bytes):
    out = np.empty((len(ct) - 16) // 8, dtype=np.float64)
    rc = lib().orc_decrypt_query(C.c_char_p(ct), C.c_int(len(ct)), C.c_char_p(key), C.c_char_p(iv), _p(out))
    return rc, out


@dataclass
class Store:
    """Flat record store + key ring: the arrays fspann_store_upload / fspann_keys_set take."""
    dim: int
    iv: np.ndarray            # uint8 [N, 12]
    ct: np.ndarray            # uint8 [N, 8*dim+16]
    key_version: np.ndarray   # int32 [N]
    keys: dict                # version -> 32-byte key
    deleted: np.ndarray | None = None  # uint8 [N]
    present: np.ndarray | None = None  # uint8 [N]

    @property
    def N(self) -> int:
        return self.iv.shape[0]

    def struct(self):
        kv = np.array(sorted(self.keys), dtype=np.int32)
        kb = np.frombuffer(b"".join(self.keys[int(v)] for v in kv), dtype=np.uint8).copy() if len(kv) else np.zeros(0, np.uint8)
        self._kv, self._kb = kv, kb
        return _StoreStruct(self.N, self.dim, self.iv.ctypes.data, self.ct.ctypes.data, self.key_version.ctypes.data,
                            self.deleted.ctypes.data if self.deleted is not None else None,
                            self.present.ctypes.data if self.present is not None else None,
                            len(kv), kv.ctypes.data if len(kv) else None, kb.ctypes.data if len(kv) else None)


def encrypt_store(vecs: np.ndarray, key_version: int, key: bytes, ivs: np.ndarray, ids: np.ndarray | None = None) -> np.ndarray:
    vecs = np.ascontiguousarray(vecs, dtype=np.float64)
    n, d = vecs.shape
    if ids is None:
        ids = np.arange(n, dtype=np.int32)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    ivs = np.ascontiguousarray(ivs, dtype=np.uint8)
    ct = np.empty((n, 8 * d + 16), dtype=np.uint8)
    rc = lib().orc_encrypt_store(_p(ids), C.c_int64(n), _p(vecs), C.c_int(d), C.c_int32(key_version), C.c_char_p(key), _p(ivs), _p(ct))
    assert rc == 0
    return ct


def refine(store: Store, q: np.ndarray, cand: np.ndarray, k: int, want_plaintext: bool = False):
    """QSI:238-322 for one query.  Returns dict(top_ids, top_dist, verdict, n_decrypted, cand_dist[, plaintext])."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    cand = np.ascontiguousarray(cand, dtype=np.int32)
    n = cand.shape[0]
    tid = np.full(k, -1, dtype=np.int32)
    tdist = np.full(k, np.nan, dtype=np.float64)
    ver = np.zeros(max(n, 1), dtype=np.uint8)
    cdist = np.full(max(n, 1), np.nan, dtype=np.float64)
    pt = np.zeros((max(n, 1), store.dim), dtype=np.float64) if want_plaintext else None
    ndec = C.c_int32(0)
    st = store.struct()
    eff = lib().orc_refine(C.byref(st), _p(q), _p(cand), C.c_int(n), C.c_int(k), _p(tid), _p(tdist), _p(ver), C.byref(ndec),
                           _p(cdist), _p(pt))
    out = dict(top_ids=tid[:eff].copy(), top_dist=tdist[:eff].copy(), verdict=ver[:n].copy(), n_decrypted=int(ndec.value),
               cand_dist=cdist[:n].copy())
    if want_plaintext:
        out["plaintext"] = pt[:n]
    return out


def search(ix: Index, store: Store, q: np.ndarray, codes: np.ndarray, k: int, probes: int, hard_cap: int, B: int,
           ham_threshold: int = 0, touched: np.ndarray | None = None):
    """QSI:100-352 for one query (with the adaptive retry).  Returns a dict describing the returned pass."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    tid = np.full(k, -1, dtype=np.int32)
    tdist = np.full(k, np.nan, dtype=np.float64)
    cid = np.full(B + 1, -1, dtype=np.int32)
    csc = np.full(B + 1, -1, dtype=np.int32)
    ver = np.zeros(B + 1, dtype=np.uint8)
    cnt = np.zeros(6, dtype=np.int64)
    ixs, sts = ix.struct(), store.struct()
    n = lib().orc_search(C.byref(ixs), C.byref(sts), _p(q), _p(codes), C.c_int(k), C.c_int(probes), C.c_int64(hard_cap),
                         C.c_int(B), C.c_int(ham_threshold), _p(tid), _p(tdist), _p(cid), _p(csc), _p(ver), _p(cnt), _p(touched))
    nc = int(cnt[5])
    return dict(top_ids=tid[:n].copy(), top_dist=tdist[:n].copy(), cand_ids=cid[:nc].copy(), cand_scores=csc[:nc].copy(),
                verdict=ver[:nc].copy(), cand_total=int(cnt[0]), cand_kept=int(cnt[1]), cand_decrypted=int(cnt[2]),
                returned=int(cnt[3]), retried=bool(cnt[4]))


def search_batch(ix: Index, store: Store, queries: np.ndarray, k: int, probes: int, hard_cap: int, B: int,
                 ham_threshold: int = 0, q0: int = 0, q1: int | None = None, out=None):
    """Sequential TokenGen->Route->Refine over queries[q0:q1] (FSA:636-747 analogue).  Thread-safe w.r.t. disjoint ranges."""
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    Q = queries.shape[0]
    if q1 is None:
        q1 = Q
    if out is None:
        out = dict(top_ids=np.full((Q, k), -1, dtype=np.int32), top_dist=np.full((Q, k), np.nan), n_ret=np.zeros(Q, dtype=np.int32),
                   counters=np.zeros((Q, 6), dtype=np.int64))
    g = ix.g
    ixs, sts = ix.struct(), store.struct()
    lib().orc_search_batch(C.byref(ixs), C.byref(sts), _p(queries), C.c_int64(q0), C.c_int64(q1), _p(g.alpha), _p(g.r), _p(g.omega),
                           C.c_int(k), C.c_int(probes), C.c_int64(hard_cap), C.c_int(B), C.c_int(ham_threshold),
                           _p(out["top_ids"]), _p(out["top_dist"]), _p(out["n_ret"]), _p(out["counters"]))
    return out


def migrate(store: Store, ids: np.ndarray, fresh_ivs: np.ndarray, target_version: int) -> int:
    """KRS:215-289 reencryptTouched, in place on the Store arrays."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    fresh_ivs = np.ascontiguousarray(fresh_ivs, dtype=np.uint8)
    st = store.struct()
    return int(lib().orc_migrate(C.c_int64(store.N), C.c_int(store.dim), _p(store.iv), _p(store.ct), _p(store.key_version), _p(ids),
                                 C.c_int64(ids.shape[0]), _p(fresh_ivs), C.c_int32(target_version), C.c_int32(st.n_keys),
                                 C.c_void_p(st.key_versions), C.c_void_p(st.keys)))


def groundtruth(base_f32: np.ndarray, queries_f32: np.ndarray, K: int):
    """GroundtruthPrecompute.run (GTP:218-276): ids int32 [Q, K] by (squared L2, id), and the squared distances."""
    base_f32 = np.ascontiguousarray(base_f32, dtype=np.float32)
    queries_f32 = np.ascontiguousarray(queries_f32, dtype=np.float32)
    n, d = base_f32.shape
    Q = queries_f32.shape[0]
    K = min(max(1, K), n)
    ids = np.empty((Q, K), dtype=np.int32)
    d2 = np.empty((Q, K), dtype=np.float64)
    lib().orc_groundtruth(_p(base_f32), C.c_int64(n), C.c_int(d), _p(queries_f32), C.c_int64(Q), C.c_int(K), _p(ids), _p(d2))
    return ids, d2


def recall_at_k(gt_row: np.ndarray, res_row: np.ndarray, n_ret: int, K: int) -> float:
    """FSA:785-794."""
    gt_row = np.ascontiguousarray(gt_row, dtype=np.int32)
    res_row = np.ascontiguousarray(res_row, dtype=np.int32)
    f = lib().orc_recall_at_k
    f.restype = C.c_double
    return float(f(_p(gt_row), _p(res_row), C.c_int(n_ret), C.c_int(K)))
