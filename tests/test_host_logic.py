"""CPU-only checks of the product's host side: the C-ABI library loads and exports every declared symbol, fails loudly
without a GPU, and the host-side Setup stand-in (hostsetup.py) and the FSP_HD kernel primitives (compiled for the host)
agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from fspann_query_system_b200 import _native, hostsetup as HS, workloads as WL
from oracle import oracle as O

CSRC = os.path.join(ROOT, "fspann_query_system_b200", "csrc")


def test_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fspann_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fspann_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for path in (_native.LIB_PATH, _native.DEBUG_LIB_PATH):
        assert os.path.exists(path), f"{path} missing: run __graft_entry__.build()"
        lib = C.CDLL(path)
        for name in declared:
            assert hasattr(lib, name), f"{os.path.basename(path)} does not export {name}"
    assert sorted(_native.EXPORTS) == declared, "the ctypes binding and include/fspann_gpu.h disagree"


def test_built_for_sm_100a_with_shared_atomics_and_no_fallback_symbols():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sym = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in sym, "the product library must not contain oracle code"


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    from fspann_query_system_b200.gpu import GpuContext
    with pytest.raises(_native.CudaError):
        GpuContext(0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "fspann_query_system_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "fspann_oracle" not in txt, f


# ---------------------------------------------------------------- hostsetup vs oracle
def test_java_hash_and_table_sizes():
    ids = np.concatenate([np.arange(0, 5000), [10 ** k for k in range(1, 10)], [123456, 999999, 2147483647]])
    assert np.array_equal(HS.java_hash_decimal(ids), np.array([O.java_hash_decimal(int(i)) for i in ids], dtype=np.uint32))
    for init, size in [(0, 0), (1, 1), (1000, 1000), (20000, 20544), (24000, 24063), (65536, 70000), (1_000_000, 1_000_000)]:
        assert HS.hashmap_final_cap(init, size) == int(O.lib().orc_hashmap_final_cap(C.c_int64(init), C.c_int64(size)))


def test_partitions_keys_kdf_encryption_match_oracle(world_factory):
    w = world_factory(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2)
    mn, mx, rep, ids = HS.build_partitions(w.codes, HS.staged_order(w.cfg.N))
    assert np.array_equal(mn, w.ix.min_key) and np.array_equal(mx, w.ix.max_key)
    assert np.array_equal(rep, w.ix.rep) and np.array_equal(ids, w.ix.ids)
    assert np.array_equal(HS.staged_order(w.cfg.N), O.staged_order(w.cfg.N))
    km = HS.KeyManager(w.master)
    for v in (1, 2, 300):
        assert km.derive(v) == O.kdf(w.master, v)
    sel = np.arange(0, 300, dtype=np.int32)
    assert np.array_equal(HS.encrypt_store(w.base[sel], sel, 1, km.derive(1), w.iv[sel]), w.ct[sel])
    q = w.queries[0]
    iv = bytes(range(12))
    assert HS.encrypt_query(q, km.derive(1), iv) == O.encrypt_query(q, km.derive(1), iv)
    assert np.array_equal(HS.decrypt_query(HS.encrypt_query(q, km.derive(1), iv), km.derive(1), iv), q)


def test_staged_order_refuses_small_index():
    with pytest.raises(RuntimeError):
        HS.staged_order(999)                 # PIS:803-808 "Cannot finalize index: only N samples collected"
    with pytest.raises(ValueError):
        O.staged_order(999)


def test_gfunction_builder_matches_oracle_to_rounding():
    """Same SplittableRandom stream and construction (Coding:184-241); log/cos come from different libms, so the arrays
    agree to a few ulp -- they are data for everything downstream."""
    cfg = WL.C1
    base = WL.base_vectors(cfg.scaled(N=1000))
    a, r, om = HS.build_gfunctions(base, 12, 2, 13, 2, 3)
    g = O.registry_init(base, 12, 2, 13, 2, 3)
    assert np.allclose(a, g.alpha, rtol=0, atol=1e-15)
    assert np.allclose(om, g.omega, rtol=1e-14) and np.allclose(r, g.r, rtol=1e-12, atol=1e-13)
    assert np.allclose(np.linalg.norm(a, axis=2), 1.0, atol=1e-12)
    assert np.all(r >= 0) and np.all(r < om)


def test_host_migrate_matches_oracle_and_changes_only_listed_records(world_factory):
    """reencryptTouched (KRS:215-289); ForwardSecurityAdversarialIT: untouched ciphertexts stay byte-identical."""
    w = world_factory(N=3000, dim=32, Q=40, T=3, D=4, m=12, lam=2)
    km = HS.KeyManager(w.master)
    km.rotate_key_only()
    iv1, ct1, kv1 = w.iv.copy(), w.ct.copy(), w.key_version.copy()
    iv2, ct2, kv2 = w.iv.copy(), w.ct.copy(), w.key_version.copy()
    ids = np.arange(0, 3000, 7, dtype=np.int32)
    fresh = WL.record_ivs(len(ids), 99)
    done = HS.migrate(iv1, ct1, kv1, ids, fresh, 2, km)
    st = O.Store(32, iv2, ct2, kv2, {1: km.derive(1), 2: km.derive(2)})
    assert O.migrate(st, ids, fresh, 2) == len(done) == len(ids)
    assert np.array_equal(iv1, iv2) and np.array_equal(ct1, ct2) and np.array_equal(kv1, kv2)
    untouched = np.setdiff1d(np.arange(3000), ids)
    assert np.array_equal(ct1[untouched], w.ct[untouched]) and np.all(kv1[untouched] == 1) and np.all(kv1[ids] == 2)
    assert HS.migrate(iv1, ct1, kv1, ids, fresh, 2, km) == []          # already upgraded -> skipped (KRS:244-245)


# ---------------------------------------------------------------- kernel primitives compiled for the host
@pytest.fixture(scope="module")
def hc():
    path = os.path.join(CSRC, "libfspann_hostcheck.so")
    assert os.path.exists(path), "run __graft_entry__.build()"
    lib = C.CDLL(path)
    lib.fsp_hc_java_hash.restype = C.c_uint32
    return lib


def test_kernel_aes_rounds_match_standard(hc):
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    rng = np.random.default_rng(1)
    # FIPS-197 appendix C.3
    key = bytes(range(32))
    out = C.create_string_buffer(16)
    hc.fsp_hc_aes_block(key, bytes.fromhex("00112233445566778899aabbccddeeff"), out)
    assert out.raw.hex() == "8ea2b7ca516745bfeafc49904b496089"
    for _ in range(50):
        key, blk = rng.bytes(32), rng.bytes(16)
        hc.fsp_hc_aes_block(key, blk, out)
        e = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
        assert out.raw == e.update(blk) + e.finalize()


def test_kernel_gf128_multiplies_match_bit_serial_reference(hc):
    rng = np.random.default_rng(2)
    for _ in range(3000):
        x = rng.integers(0, 2 ** 64, size=2, dtype=np.uint64)
        y = rng.integers(0, 2 ** 64, size=2, dtype=np.uint64)
        zf, zr = np.zeros(2, np.uint64), np.zeros(2, np.uint64)
        hc.fsp_hc_gfmul(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), zf.ctypes.data_as(C.c_void_p), zr.ctypes.data_as(C.c_void_p))
        assert np.array_equal(zf, zr)
    h = rng.integers(0, 2 ** 64, size=2, dtype=np.uint64)
    xs = rng.integers(0, 2 ** 64, size=4000, dtype=np.uint64)
    assert hc.fsp_hc_shoup_check(h.ctypes.data_as(C.c_void_p), xs.ctypes.data_as(C.c_void_p), 2000) == 0


@pytest.mark.parametrize("dim", [128, 100, 96, 33, 7, 1])
def test_kernel_record_decrypt_matches_oracle(hc, dim):
    rng = np.random.default_rng(dim)
    for id_ in (0, 9, 999, 1000, 123456, 2147483647):
        for ver in (1, 12, 345):
            key, iv, v = rng.bytes(32), rng.bytes(12), rng.normal(size=dim)
            ct = O.encrypt_point(id_, ver, v, key, iv)
            pl = C.create_string_buffer(8 * dim)
            assert hc.fsp_hc_decrypt_record(key, iv, C.c_int64(id_), ver, dim, ct, pl) == 1
            assert np.array_equal(np.frombuffer(pl.raw, dtype=">f8"), v)
            bad = bytearray(ct)
            bad[rng.integers(0, len(ct))] ^= 0x40
            assert hc.fsp_hc_decrypt_record(key, iv, C.c_int64(id_), ver, dim, bytes(bad), pl) == 0
            assert hc.fsp_hc_decrypt_record(key, iv, C.c_int64(id_), ver + 1, dim, ct, pl) == 0
            buf = C.create_string_buffer(48)
            n = hc.fsp_hc_aad(C.c_int64(id_), ver, dim, buf)
            assert buf.raw[:n] == O.aad(id_, ver, dim)


def test_kernel_java_hash(hc):
    for i in list(range(3000)) + [10 ** k for k in range(1, 10)] + [2147483647]:
        assert hc.fsp_hc_java_hash(i) == O.java_hash_decimal(i)
