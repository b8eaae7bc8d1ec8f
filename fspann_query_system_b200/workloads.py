"""Synthetic workloads for BASELINE.json's configs.

Values are drawn as float32 and widened to float64 exactly like the reference's loaders do
(loader/src/main/java/com/fspann/loader/FvecsLoader.java:27-30).  Pure numpy; no oracle, no CUDA.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import numpy as np

MASTER_KEY = hashlib.sha256(b"fspann-b200-bench").digest()


@dataclass(frozen=True)
class Config:
    name: str
    N: int
    dim: int
    Q: int
    T: int
    D: int
    m: int
    lam: int
    seed: int          # paper.seed (config_sift1m.json:14)
    probes: int        # DEFAULT_MAX_PROBES PIS:93
    B: int             # runtime.refinementLimit
    k: int
    max_global: int    # runtime.maxGlobalCandidates
    shape: str         # "sift" | "glove" | "deep"
    base_seed: int
    query_seed: int
    centres: int = 0

    @property
    def hard_cap(self) -> int:  # PIS:612-615
        return max(self.max_global, self.B)

    @property
    def W(self) -> int:
        return (self.m * self.lam + 63) // 64

    def scaled(self, N: int | None = None, Q: int | None = None, name: str | None = None) -> "Config":
        d = dict(self.__dict__)
        if N is not None:
            d["N"] = N
        if Q is not None:
            d["Q"] = Q
        d["name"] = name or f"{self.name}[N={d['N']},Q={d['Q']}]"
        return Config(**d)


C1 = Config("C1-sift10k", 10_000, 128, 100, 4, 8, 24, 2, 13, 5, 256, 10, 20_000, "sift", 1001, 2001, 256)
C2 = Config("C2-sift1m", 1_000_000, 128, 10_000, 8, 8, 24, 2, 13, 5, 1024, 10, 24_000, "sift", 1002, 2002, 4096)
C3 = Config("C3-glove1.2m", 1_200_000, 100, 10_000, 8, 8, 22, 2, 13, 5, 1024, 10, 24_000, "glove", 1003, 2003)
C4 = Config("C4-deep100m", 100_000_000, 96, 10_000, 8, 8, 24, 2, 13, 5, 1024, 10, 24_000, "deep", 1004, 2004)
# config 4's shape (Deep-like 96-d, T = D = 8) at a size one GPU's host can build in seconds: the bench's single-GPU stand-in for C4
C4S = Config("C4s-deep4m", 4_000_000, 96, 10_000, 8, 8, 24, 2, 13, 5, 1024, 10, 24_000, "deep", 1004, 2004)
# the reference's only published operating point (README.md:300, ART 2828 ms): profile SIFT_P6_BALANCED of config_sift1m.json:59-71 --
# tables 6, probeOverride 6, refinementLimit 16000, maxGlobalCandidates 20000 -- queried at MAX_K = 100 like FSA.runQueries (FSA:634)
P6 = Config("P6-sift1m-balanced", 1_000_000, 128, 1_000, 6, 8, 24, 2, 13, 6, 16_000, 100, 20_000, "sift", 1002, 2002, 4096)
# a profile of the reference whose HARD_CAP binds (sift1m_sub1.json, profile 2: tables 5, divisions 12, m 22, probeOverride 6, refinementLimit 6000,
# maxGlobalCandidates 8000: 60 x 6 x 64 = 23 040 positions per query against a cap of 8 000), at full SIFT1M shape
SUB1 = Config("SUB1-sift1m-sub1-p2", 1_000_000, 128, 1_000, 5, 12, 22, 2, 13, 6, 6_000, 100, 8_000, "sift", 1002, 2002, 4096)
CONFIGS = {"C1": C1, "C2": C2, "C3": C3, "C4": C4, "C4s": C4S, "C5": C2, "P6": P6, "SUB1": SUB1}


def _sift(n: int, dim: int, centres: int, seed: int, centre_seed: int) -> np.ndarray:
    crng = np.random.Generator(np.random.PCG64(centre_seed))
    cen = crng.uniform(0.0, 128.0, size=(centres, dim)).astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n, dim), dtype=np.float32)
    step = 262_144
    for s in range(0, n, step):
        e = min(n, s + step)
        a = rng.integers(0, centres, size=e - s)
        x = cen[a] + rng.normal(0.0, 20.0, size=(e - s, dim)).astype(np.float32)
        out[s:e] = np.rint(np.clip(x, 0.0, 255.0))
    return out


def base_vectors(cfg: Config, lo: int = 0, hi: int | None = None) -> np.ndarray:
    """Rows [lo, hi) of the base set as float64 (float32 values widened)."""
    hi = cfg.N if hi is None else hi
    if cfg.shape == "sift":
        # cluster assignment stream is sequential, so generate from 0 and slice (N <= a few M for this shape)
        return _sift(hi, cfg.dim, cfg.centres, cfg.base_seed, cfg.base_seed + 7)[lo:hi].astype(np.float64)
    chunk = 1_000_000
    parts = []
    for c in range(lo // chunk, (hi - 1) // chunk + 1):
        rng = np.random.Generator(np.random.PCG64(cfg.base_seed + c))
        if cfg.shape == "glove":
            x = rng.normal(0.0, 0.4, size=(chunk, cfg.dim)).astype(np.float32)
        else:
            x = rng.normal(0.0, 1.0, size=(chunk, cfg.dim)).astype(np.float32)
            x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
        a, b = max(lo, c * chunk) - c * chunk, min(hi, (c + 1) * chunk) - c * chunk
        parts.append(x[a:b])
    return np.concatenate(parts).astype(np.float64)


def query_vectors(cfg: Config, Q: int | None = None) -> np.ndarray:
    Q = cfg.Q if Q is None else Q
    if cfg.shape == "sift":
        return _sift(Q, cfg.dim, cfg.centres, cfg.query_seed, cfg.base_seed + 7).astype(np.float64)
    rng = np.random.Generator(np.random.PCG64(cfg.query_seed))
    if cfg.shape == "glove":
        return rng.normal(0.0, 0.4, size=(Q, cfg.dim)).astype(np.float32).astype(np.float64)
    x = rng.normal(0.0, 1.0, size=(Q, cfg.dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return x.astype(np.float64)


def record_ivs(n: int, seed: int) -> np.ndarray:
    """Deterministic stand-in for SecureRandom 96-bit IVs (AGC:66-67): uint8 [n, 12]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(n, 12), dtype=np.uint8)
