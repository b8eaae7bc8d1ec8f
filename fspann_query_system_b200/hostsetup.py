"""Host-side Setup stand-in for the Java host (Setup/Rotate/Migrate stay on the host, north_star).

Product-side host logic (no oracle, no CUDA): what the reference's Java code does before the query hot path starts,
restated in numpy so the GPU library can be fed exactly what a Java host would feed it over the C ABI:
  * GreedyPartitioner.build for every (table, division)             index/.../paper/GreedyPartitioner.java:37-76
    (incl. java.util.HashMap iteration order of the staged ids,     index/.../paper/PartitionedIndexService.java:412-425)
  * staged order of vectors (999.., then the 999 parked ones)       PartitionedIndexService.java:280-298, 821-831
  * KeyManager.deriveSessionKey = HMAC-SHA256(K_M, be32(version))   keymanagement/.../KeyManager.java:221-237
  * AesGcmCryptoService.encryptToPoint / encryptQuery               crypto/.../AesGcmCryptoService.java:55-112,169-186
  * KeyRotationServiceImpl.reencryptTouched (Migrate)               keymanagement/.../KeyRotationServiceImpl.java:215-289
"""
from __future__ import annotations

import hashlib
import hmac
from dataclasses import dataclass

import numpy as np
from cryptography.exceptions import InvalidTag
from cryptography.hazmat.primitives.ciphers.aead import AESGCM

BLOCK = 64            # DEFAULT_GREEDY_BLOCK_SIZE (PIS:92)
MIN_SAMPLE_SIZE = 1000  # PIS:50


def java_hash_decimal(ids: np.ndarray) -> np.ndarray:
    """HashMap.hash(String.valueOf(id)) for non-negative ids, vectorised (uint32)."""
    ids = np.asarray(ids, dtype=np.int64)
    ndig = np.ones(ids.shape, dtype=np.int64)
    t = ids // 10
    while np.any(t > 0):
        ndig += t > 0
        t //= 10
    h = np.zeros(ids.shape, dtype=np.uint64)
    maxd = int(ndig.max()) if ids.size else 1
    for pos in range(maxd):           # most-significant digit first
        exp = ndig - 1 - pos
        active = exp >= 0
        digit = (ids // np.power(10, np.maximum(exp, 0))) % 10
        h = np.where(active, (h * np.uint64(31) + np.uint64(48) + digit.astype(np.uint64)) & np.uint64(0xFFFFFFFF), h)
    h = h.astype(np.uint32)
    return h ^ (h >> np.uint32(16))


def table_size_for(cap: int) -> int:
    n = 1
    while n < cap and n < (1 << 30):
        n <<= 1
    return n


def hashmap_final_cap(initial_capacity: int, size: int) -> int:
    cap = table_size_for(initial_capacity) if initial_capacity > 0 else 1
    while size > 0.75 * cap and cap < (1 << 30):
        cap <<= 1
    return cap


def staged_order(N: int) -> np.ndarray:
    """Insertion order into PIS's per-table staging lists: the registry initialises when the 1000th vector arrives
    (PIS:280-290), that vector and all later ones are staged immediately, the first 999 at finalize (PIS:821-831)."""
    if N < MIN_SAMPLE_SIZE:
        raise RuntimeError(f"Cannot finalize index: only {N} samples collected (< MIN_SAMPLE_SIZE)")  # PIS:803-808
    return np.concatenate([np.arange(MIN_SAMPLE_SIZE - 1, N, dtype=np.int32), np.arange(0, MIN_SAMPLE_SIZE - 1, dtype=np.int32)])


def compute_keys(codes_w0: np.ndarray) -> np.ndarray:
    """GreedyPartitioner.computeKey (GP:87-96): code bit i (i < 63) -> key bit 62-i."""
    x = np.ascontiguousarray(codes_w0, dtype=np.uint64)
    # bit-reverse 64 bits, then drop the (reversed) bit 63 of the code
    x = ((x >> np.uint64(1)) & np.uint64(0x5555555555555555)) | ((x & np.uint64(0x5555555555555555)) << np.uint64(1))
    x = ((x >> np.uint64(2)) & np.uint64(0x3333333333333333)) | ((x & np.uint64(0x3333333333333333)) << np.uint64(2))
    x = ((x >> np.uint64(4)) & np.uint64(0x0F0F0F0F0F0F0F0F)) | ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4))
    x = x.byteswap()
    return (x >> np.uint64(1)).astype(np.int64)


@dataclass
class RoutingIndex:
    dim: int
    T: int
    D: int
    m: int
    lam: int
    alpha: np.ndarray   # [T*D, m, dim]
    r: np.ndarray       # [T*D, m]
    omega: np.ndarray   # [T*D, m]
    N: int
    min_key: np.ndarray  # int64 [T*D, P]
    max_key: np.ndarray  # int64 [T*D, P]
    rep: np.ndarray      # uint64 [T*D, P, W]
    ids: np.ndarray      # int32 [T*D, N]


def build_partitions(codes_by_id: np.ndarray, staged_ids: np.ndarray):
    """GP.build for every (t,d).  codes_by_id uint64 [N, TD, W]; staged_ids = HashMap insertion order.
    Returns (min_key[TD,P], max_key[TD,P], rep[TD,P,W], ids[TD,N])."""
    staged_ids = np.asarray(staged_ids, dtype=np.int32)
    N = staged_ids.shape[0]
    _, TD, W = codes_by_id.shape
    P = (N + BLOCK - 1) // BLOCK
    cap = hashmap_final_cap(N, N)                       # new HashMap<>(S.staged.size()) (PIS:413)
    bucket = java_hash_decimal(staged_ids) & np.uint32(cap - 1)
    it_order = np.argsort(bucket, kind="stable")        # HashMap iteration: bucket asc, insertion order inside
    ids_it = staged_ids[it_order]
    mn = np.zeros((TD, P), dtype=np.int64)
    mx = np.zeros((TD, P), dtype=np.int64)
    rep = np.zeros((TD, P, W), dtype=np.uint64)
    ids = np.zeros((TD, N), dtype=np.int32)
    starts = np.arange(0, N, BLOCK)
    ends = np.minimum(starts + BLOCK, N)
    mids = starts + ((ends - starts - 1) >> 1)          # GP:60
    for td in range(TD):
        c = codes_by_id[ids_it, td, :]                  # codes in iteration order
        keys = compute_keys(c[:, 0])
        o = np.argsort(keys, kind="stable")             # GP:51 (List.sort is stable)
        ks = keys[o]
        ids[td] = ids_it[o]
        mn[td] = ks[starts]
        mx[td] = ks[ends - 1]
        rep[td] = c[o[mids]]
    return mn, mx, rep, ids


# ---------------------------------------------------------------- keys & crypto (host side of Setup / Rotate / Migrate)
class KeyManager:
    """KeyManager / KeyRotationServiceImpl stand-in: K_v = HMAC-SHA256(K_M, be32(v))[:32] (KM:221-237)."""

    def __init__(self, master_key: bytes):
        assert len(master_key) == 32
        self.master = master_key
        self.current = 1
        self.live = {1}

    def derive(self, version: int) -> bytes:
        return hmac.new(self.master, int(version).to_bytes(4, "big", signed=True), hashlib.sha256).digest()[:32]

    def get_version(self, version: int) -> bytes:           # KRS:82-88
        if version not in self.live:
            raise ValueError(f"Unknown key version: {version}")
        return self.derive(version)

    def rotate_key_only(self) -> int:                       # KRS:292-298
        self.current += 1
        self.live.add(self.current)
        return self.current

    def retire(self, version: int):                         # KM:274-317 (caller checks live counts)
        self.live.discard(version)


def aad_bytes(id_: int, key_version: int, dim: int) -> bytes:   # EP:80-83
    return b"id:%d|v:%d|d:%d" % (id_, key_version, dim)


def encrypt_store(vecs: np.ndarray, ids: np.ndarray, key_version: int, key: bytes, ivs: np.ndarray) -> np.ndarray:
    """encryptToPoint for many records (AGC:55-112): returns uint8 [n, 8*dim+16]."""
    n, dim = vecs.shape
    be = np.ascontiguousarray(vecs, dtype=">f8")
    a = AESGCM(key)
    out = np.empty((n, 8 * dim + 16), dtype=np.uint8)
    ivb = np.ascontiguousarray(ivs, dtype=np.uint8).tobytes()
    row = 8 * dim
    raw = be.tobytes()
    for i in range(n):
        ct = a.encrypt(ivb[12 * i:12 * i + 12], raw[row * i:row * (i + 1)], aad_bytes(int(ids[i]), key_version, dim))
        out[i] = np.frombuffer(ct, dtype=np.uint8)
    return out


def encrypt_query(vec: np.ndarray, key: bytes, iv: bytes) -> bytes:     # AGC:169-186 (no AAD)
    return AESGCM(key).encrypt(iv, np.ascontiguousarray(vec, dtype=">f8").tobytes(), None)


def decrypt_query(ct: bytes, key: bytes, iv: bytes) -> np.ndarray:      # AGC:189-204
    try:
        pt = AESGCM(key).decrypt(iv, ct, None)
    except InvalidTag as e:
        raise RuntimeError("Query decryption failed") from e
    return np.frombuffer(pt, dtype=">f8").astype(np.float64)


def migrate(iv: np.ndarray, ct: np.ndarray, key_version: np.ndarray, ids, fresh_ivs: np.ndarray, target_version: int,
            keys: KeyManager):
    """reencryptTouched (KRS:215-289), in place on host arrays.  Returns the list of ids actually re-encrypted."""
    dim = (ct.shape[1] - 16) // 8
    tkey = AESGCM(keys.get_version(target_version))
    done = []
    for j, id_ in enumerate(ids):
        id_ = int(id_)
        old = int(key_version[id_])
        if old >= target_version:
            continue
        try:
            okey = AESGCM(keys.get_version(old))
            pt = okey.decrypt(iv[id_].tobytes(), ct[id_].tobytes(), aad_bytes(id_, old, dim))
        except (ValueError, InvalidTag):
            continue   # forward-secure skip (KRS:277-279)
        iv[id_] = fresh_ivs[j]
        ct[id_] = np.frombuffer(tkey.encrypt(iv[id_].tobytes(), pt, aad_bytes(id_, target_version, dim)), dtype=np.uint8)
        key_version[id_] = target_version
        done.append(id_)
    return done


# ---------------------------------------------------------------- GFunction construction (Setup, host side)
_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _splittable_doubles(seed: int, n: int) -> np.ndarray:
    """First n values of new SplittableRandom(seed).nextDouble() (SplitMix64, default gamma)."""
    with np.errstate(over="ignore"):
        k = np.arange(1, n + 1, dtype=np.uint64)
        state = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + k * _GAMMA
        bits = _mix64(state)
    return (bits >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def build_gfunctions(sample: np.ndarray, m: int, lam: int, base_seed: int, T: int, D: int):
    """GFunctionRegistry.initialize -> Coding.buildFromSample per (table, division) (GFR:63-147, Coding:184-241):
    unit-norm Gaussian rows (Box-Muller over SplittableRandom, Coding:342-347), omega = max(1e-6, range)/2.5 over the
    sample's projections (sequential FP64 dot, Coding:349-353), r ~ U[0, omega) drawn after all of alpha.
    The arrays are DATA for everything downstream (a Java host would upload its own)."""
    sample = np.ascontiguousarray(sample, dtype=np.float64)
    n, d = sample.shape
    alpha = np.empty((T * D, m, d))
    r = np.empty((T * D, m))
    omega = np.empty((T * D, m))
    for t in range(T):
        for dv in range(D):
            g = t * D + dv
            seed = base_seed + t * 1_000_003 + dv                   # GFR:291-293
            u = _splittable_doubles(seed, 2 * m * d + m)
            u1 = np.maximum(np.float64(4.9e-324), u[0:2 * m * d:2])
            u2 = u[1:2 * m * d:2]
            a = (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).reshape(m, d)
            norm = np.zeros(m)
            for i in range(d):
                norm = norm + a[:, i] * a[:, i]
            a = a / np.sqrt(np.maximum(1e-12, norm))[:, None]
            y = np.zeros((n, m))
            for i in range(d):                                      # index order, mul then add: Java's dot()
                y = y + sample[:, i, None] * a[None, :, i]
            rng_ = np.maximum(1e-6, y.max(axis=0) - y.min(axis=0))
            om = rng_ / 2.5
            om = np.where(om > 0, om, 1e-3)
            alpha[g], omega[g] = a, om
            r[g] = u[2 * m * d:] * om
    return alpha, r, omega
