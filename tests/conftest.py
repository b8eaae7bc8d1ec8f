import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class World:
    """A small FSPANN deployment built with the ORACLE (the checker): GFunctions, routing index, encrypted store."""

    def __init__(self, N, dim, Q, T, D, m, lam, seed=13, shape="sift", data_seed=7, n_versions=1):
        from oracle import oracle as O
        from fspann_query_system_b200 import workloads as WL
        cfg = WL.Config(f"t-{shape}-{N}x{dim}", N, dim, Q, T, D, m, lam, seed, 5, 64, 10, 20000, shape, data_seed, data_seed + 1000,
                        centres=64)
        self.cfg = cfg
        self.base = WL.base_vectors(cfg)
        self.queries = WL.query_vectors(cfg)
        self.g = O.registry_init(self.base[:1000], m, lam, seed, T, D)
        self.codes = O.tokengen_batch(self.base, self.g)
        self.ix = O.index_build(self.codes, self.g, O.staged_order(N))
        assert self.ix.max_chain < 9
        self.master = WL.MASTER_KEY
        self.keys = {v: O.kdf(self.master, v) for v in range(1, n_versions + 1)}
        self.iv = WL.record_ivs(N, data_seed + 5)
        self.key_version = np.ones(N, dtype=np.int32)
        if n_versions > 1:
            rng = np.random.default_rng(data_seed + 9)
            self.key_version = rng.integers(1, n_versions + 1, size=N).astype(np.int32)
        self.ct = np.empty((N, 8 * dim + 16), dtype=np.uint8)
        for v in self.keys:
            sel = np.nonzero(self.key_version == v)[0].astype(np.int32)
            if len(sel):
                self.ct[sel] = O.encrypt_store(self.base[sel], v, self.keys[v], self.iv[sel], ids=sel)
        self.store = O.Store(dim, self.iv, self.ct, self.key_version, dict(self.keys))

    def gpu_context(self, debug=False):
        from fspann_query_system_b200.gpu import GpuContext
        g, ix = self.g, self.ix
        ctx = GpuContext(0, debug=debug)
        ctx.set_option("route_small_v1", 0)       # test batches are smaller than the SM count: keep Route on the two-CTA kernel (the default for real batches)
        ctx.routing_upload(g.dim, g.T, g.D, g.m, g.lam, g.alpha, g.r, g.omega, ix.min_key, ix.max_key, ix.rep, ix.ids)
        for v, k in self.store.keys.items():
            ctx.keys_set(v, k)
        ctx.store_upload(g.dim, self.store.iv, self.store.ct, self.store.key_version)
        return ctx


_worlds = {}


@pytest.fixture(scope="session")
def world_factory():
    def make(**kw):
        key = tuple(sorted(kw.items()))
        if key not in _worlds:
            _worlds[key] = World(**kw)
        return _worlds[key]
    return make
