"""reader of the reference's persisted records (.point files = Java serialization of EncryptedPoint, RocksDB metadata
value strings).  The fixtures come from a serializer written here from the Java Object Serialization Specification (no JVM in this
image, no .point file in the reference tree), emitting what ObjectOutputStream.writeObject(EncryptedPoint) emits: class descriptor
with the fields in canonical order (primitives by name, then objects by name), List.of() as java.util.CollSer, the second empty
list as a back-reference."""
import os
import struct

import numpy as np
import pytest

from fspann_query_system_b200 import javaio as J


def _utf(s):
    b = s.encode()
    return struct.pack(">H", len(b)) + b


def serialize_encrypted_point(id_, version, iv, ct, key_version, dim, shard=0, lists="of"):
    """ObjectOutputStream output for new EncryptedPoint(id, version, iv, ct, keyVersion, dim, shard, List.of(), List.of())."""
    out = bytearray(struct.pack(">HH", J.STREAM_MAGIC, J.STREAM_VERSION))
    handles = 0

    def new_handle():
        nonlocal handles
        handles += 1
        return J.BASE_HANDLE + handles - 1
    out += bytes([J.TC_OBJECT, J.TC_CLASSDESC]) + _utf("com.fspann.common.EncryptedPoint") + struct.pack(">q", 1)
    new_handle()                                                           # class descriptor
    out += bytes([J.SC_SERIALIZABLE]) + struct.pack(">H", 9)
    for name in ("dimension", "keyVersion", "shardId", "version"):          # primitive fields, sorted by name
        out += b"I" + _utf(name)
    h_list = h_bytes = None
    for t, name, sig in (("L", "buckets", "Ljava/util/List;"), ("[", "ciphertext", "[B"), ("L", "id", "Ljava/lang/String;"), ("[", "iv", "[B"),
                         ("L", "metadata", "Ljava/util/List;")):
        out += t.encode() + _utf(name)
        if sig == "Ljava/util/List;" and h_list is not None:
            out += bytes([J.TC_REFERENCE]) + struct.pack(">I", h_list)      # type strings are shared objects
        elif sig == "[B" and h_bytes is not None:
            out += bytes([J.TC_REFERENCE]) + struct.pack(">I", h_bytes)
        else:
            out += bytes([J.TC_STRING]) + _utf(sig)
            h = new_handle()
            if sig == "Ljava/util/List;":
                h_list = h
            elif sig == "[B":
                h_bytes = h
    out += bytes([J.TC_ENDBLOCKDATA, J.TC_NULL])                           # no class annotation, no serializable superclass
    new_handle()                                                           # the object itself
    out += struct.pack(">iiii", dim, key_version, shard, version)
    # buckets
    if lists == "of":                                                      # List.of() -> writeReplace -> java.util.CollSer(IMM_LIST)
        out += bytes([J.TC_OBJECT, J.TC_CLASSDESC]) + _utf("java.util.CollSer") + struct.pack(">q", 6309168927139932177)
        new_handle()
        out += bytes([J.SC_SERIALIZABLE | J.SC_WRITE_METHOD]) + struct.pack(">H", 1) + b"I" + _utf("tag") + bytes([J.TC_ENDBLOCKDATA, J.TC_NULL])
        h_empty = new_handle()
        out += struct.pack(">i", 1) + bytes([J.TC_BLOCKDATA, 4]) + struct.pack(">i", 0) + bytes([J.TC_ENDBLOCKDATA])
    else:                                                                  # new ArrayList<>() variant
        out += bytes([J.TC_OBJECT, J.TC_CLASSDESC]) + _utf("java.util.ArrayList") + struct.pack(">q", 8683452581122892189)
        new_handle()
        out += bytes([J.SC_SERIALIZABLE | J.SC_WRITE_METHOD]) + struct.pack(">H", 1) + b"I" + _utf("size") + bytes([J.TC_ENDBLOCKDATA, J.TC_NULL])
        h_empty = new_handle()
        out += struct.pack(">i", 0) + bytes([J.TC_BLOCKDATA, 4]) + struct.pack(">i", 0) + bytes([J.TC_ENDBLOCKDATA])
    # ciphertext
    out += bytes([J.TC_ARRAY, J.TC_CLASSDESC]) + _utf("[B") + struct.pack(">q", -5984413125824719648)
    h_bcls = new_handle()
    out += bytes([J.SC_SERIALIZABLE]) + struct.pack(">H", 0) + bytes([J.TC_ENDBLOCKDATA, J.TC_NULL])
    new_handle()
    out += struct.pack(">i", len(ct)) + bytes(ct)
    # id
    out += bytes([J.TC_STRING]) + _utf(id_)
    new_handle()
    # iv: same array class -> back-reference to its descriptor
    out += bytes([J.TC_ARRAY, J.TC_REFERENCE]) + struct.pack(">I", h_bcls)
    new_handle()
    out += struct.pack(">i", len(iv)) + bytes(iv)
    # metadata: the same List.of() singleton -> back-reference to the replacement object
    out += bytes([J.TC_REFERENCE]) + struct.pack(">I", h_empty)
    return bytes(out)


@pytest.mark.parametrize("lists", ["of", "arraylist"])
def test_parse_encrypted_point_round_trip(lists):
    rng = np.random.default_rng(1)
    dim = 24
    iv, ct = rng.integers(0, 256, 12, dtype=np.uint8).tobytes(), rng.integers(0, 256, 8 * dim + 16, dtype=np.uint8).tobytes()
    rec = J.parse_encrypted_point(serialize_encrypted_point("123456", 3, iv, ct, 3, dim, shard=2, lists=lists))
    assert (rec.id, rec.version, rec.key_version, rec.dimension, rec.shard_id) == ("123456", 3, 3, dim, 2)
    assert rec.iv == iv and rec.ciphertext == ct


def test_parser_rejects_foreign_and_damaged_streams():
    good = serialize_encrypted_point("7", 1, bytes(12), bytes(8 * 4 + 16), 1, 4)
    with pytest.raises(J.JavaStreamError):
        J.parse_encrypted_point(b"\x00\x01" + good[2:])                    # wrong magic
    with pytest.raises(J.JavaStreamError):
        J.parse_encrypted_point(good[:-20])                                # truncated
    with pytest.raises(J.JavaStreamError):
        J.parse_encrypted_point(good.replace(b"com.fspann.common.EncryptedPoint", b"com.fspann.common.EncryptedPoinX"))
    with pytest.raises(J.JavaStreamError):
        J.parse_encrypted_point(serialize_encrypted_point("7", 1, bytes(11), bytes(8 * 4 + 16), 1, 4))   # IV not 96 bits


def test_vector_metadata_string():
    assert J.parse_vector_metadata("version=3;shardId=0;dim=128") == {"version": "3", "shardId": "0", "dim": "128"}
    assert J.parse_vector_metadata("a\\=b=c\\;d;e=f") == {"a=b": "c;d", "e": "f"}                      # RDB:815-821 escapes
    assert J.parse_vector_metadata("") == {}


def test_scan_points_dir_matches_the_store_after_a_migrate(tmp_path):
    """A persisted deployment after Rotate + partial Migrate: v1 files for everything, v2 files for the migrated ids (old files not
    yet cleaned up).  The scan returns exactly the arrays fspann_store_upload takes; with RocksDB metadata it resolves like
    loadEncryptedPoint, without it the highest version wins; ids whose file is missing or unreadable are reported absent."""
    from fspann_query_system_b200 import hostsetup as HS, workloads as WL
    rng = np.random.default_rng(3)
    n, dim = 40, 8
    base = rng.normal(0, 1, size=(n, dim))
    km = HS.KeyManager(WL.MASTER_KEY)
    iv1 = WL.record_ivs(n, 11)
    ct1 = HS.encrypt_store(base, np.arange(n, dtype=np.int32), 1, km.derive(1), iv1)
    ivs, cts, ver = iv1.copy(), ct1.copy(), np.ones(n, dtype=np.int32)
    km.rotate_key_only()
    mig = list(range(0, n, 3))
    HS.migrate(ivs, cts, ver, mig, WL.record_ivs(len(mig), 12), 2, km)
    for i in range(n):
        os.makedirs(tmp_path / "v1", exist_ok=True)
        (tmp_path / "v1" / f"{i}.point").write_bytes(serialize_encrypted_point(str(i), 1, iv1[i].tobytes(), ct1[i].tobytes(), 1, dim))
    os.makedirs(tmp_path / "v2", exist_ok=True)
    for i in mig:
        (tmp_path / "v2" / f"{i}.point").write_bytes(serialize_encrypted_point(str(i), 2, ivs[i].tobytes(), cts[i].tobytes(), 2, dim))
    (tmp_path / "v2" / "notes.txt").write_text("ignored")
    iv, ct, kv, present, d = J.scan_points_dir(str(tmp_path))
    assert d == dim and present.all()
    assert np.array_equal(iv, ivs) and np.array_equal(ct, cts) and np.array_equal(kv, ver)
    # metadata-driven resolution: id 3 still points at v1 (its Migrate had not committed), id 5 has no metadata, id 6's file is gone
    meta = {str(i): f"version={ver[i]};shardId=0;dim={dim}" for i in range(n)}
    meta["3"] = f"version=v1;shardId=0;dim={dim}"
    del meta["5"]
    os.remove(tmp_path / "v2" / "6.point")
    iv_m, ct_m, kv_m, present_m, _ = J.scan_points_dir(str(tmp_path), meta)
    assert not present_m[5] and not present_m[6] and present_m.sum() == n - 2
    assert kv_m[3] == 1 and np.array_equal(ct_m[3], ct1[3]) and np.array_equal(iv_m[3], iv1[3])
    ok = present_m & (np.arange(n) != 3)
    assert np.array_equal(ct_m[ok], cts[ok]) and np.array_equal(kv_m[ok], ver[ok])
    # a damaged file is skipped like loadPointIfActive swallows the exception
    (tmp_path / "v1" / "1.point").write_bytes(b"garbage")
    assert not J.scan_points_dir(str(tmp_path))[3][1]
    # the scanned records decrypt with the oracle (AAD binds id / key version / dim)
    from oracle import oracle as O
    for i in (0, 1 + 1, 9):
        rc, pt = O.decrypt_point(i, int(kv[i]), dim, km.derive(int(kv[i])), iv[i].tobytes(), ct[i].tobytes())
        assert rc == 0 and np.array_equal(pt, base[i])
