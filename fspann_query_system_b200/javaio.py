"""Host-side reader for the reference's persisted records, so an existing FSPANN deployment can be mirrored into the
HBM store without re-encrypting.

What the reference writes (common/.../RocksDBMetadataManager.java:342-375, 530-544; common/.../PersistenceUtils.java:23-47):
  * <baseDir>/v<version>/<id>.point  -- `new ObjectOutputStream(...).writeObject(EncryptedPoint)`: the Java Object Serialization
    Stream Protocol (magic 0xACED, version 5) of common/.../EncryptedPoint.java:15-27 (id, version, iv, ciphertext, keyVersion,
    dimension, shardId, buckets, metadata);
  * RocksDB key <id> -> "version=<v>;shardId=<s>;dim=<d>" with '=' and ';' escaped by a backslash (RDB:736-752, 815-821).
RocksDB's own SST/WAL files need the RocksDB library, which this image does not have: `parse_vector_metadata` takes the VALUE strings
(e.g. from `ldb scan`), and `scan_points_dir` works from the .point files alone -- for every id it takes the file in the highest
v<version> directory, which is the one loadEncryptedPoint resolves through the metadata (a Migrate writes the new version's file and
only queues the older ones for cleanup, RDB:373).

PARITY STATUS: the parser follows the published stream grammar (Java Object Serialization Specification, ch. 6); no real .point file
exists in the reference tree and no JVM is available here, so the fixtures in tests/ are produced by a serializer written from the same
specification -- unpinned against a real JVM.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass

import numpy as np

STREAM_MAGIC, STREAM_VERSION = 0xACED, 5
TC_NULL, TC_REFERENCE, TC_CLASSDESC, TC_OBJECT, TC_STRING, TC_ARRAY, TC_CLASS, TC_BLOCKDATA, TC_ENDBLOCKDATA = 0x70, 0x71, 0x72, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78
TC_RESET, TC_BLOCKDATALONG, TC_EXCEPTION, TC_LONGSTRING, TC_PROXYCLASSDESC, TC_ENUM = 0x79, 0x7A, 0x7B, 0x7C, 0x7D, 0x7E
BASE_HANDLE = 0x7E0000
SC_WRITE_METHOD, SC_SERIALIZABLE, SC_EXTERNALIZABLE, SC_BLOCK_DATA, SC_ENUM = 0x01, 0x02, 0x04, 0x08, 0x10
_PRIM = {"B": ">b", "C": ">H", "D": ">d", "F": ">f", "I": ">i", "J": ">q", "S": ">h", "Z": ">?"}


class JavaStreamError(ValueError):
    pass


@dataclass
class ClassDesc:
    name: str
    uid: int
    flags: int
    fields: list          # [(typecode, name, class_name or None)]
    super: "ClassDesc | None"


@dataclass
class JavaObject:
    cls: ClassDesc
    fields: dict          # field name -> value, over the whole class hierarchy
    annotations: list     # objects / raw block data written by writeObject methods


class JavaObjectStream:
    """Minimal reader of the serialization stream grammar: objects, class descriptors, strings, arrays, references, block data."""

    def __init__(self, data: bytes):
        self.b, self.p, self.handles = data, 0, []
        magic, ver = struct.unpack_from(">HH", data, 0)
        if magic != STREAM_MAGIC or ver != STREAM_VERSION:
            raise JavaStreamError("not a Java serialization stream")
        self.p = 4

    def _take(self, n):
        if self.p + n > len(self.b):
            raise JavaStreamError("truncated stream")
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def _u(self, fmt):
        return struct.unpack(fmt, self._take(struct.calcsize(fmt)))[0]

    def _utf(self, long=False):
        n = self._u(">q") if long else self._u(">H")
        return self._take(n).decode("utf-8", errors="surrogatepass")     # modified UTF-8 == UTF-8 for the ASCII this path writes

    def _new_handle(self, obj):
        self.handles.append(obj)
        return len(self.handles) - 1

    def read_class_desc(self):
        tc = self._u(">B")
        if tc == TC_NULL:
            return None
        if tc == TC_REFERENCE:
            return self._ref()
        if tc == TC_PROXYCLASSDESC:
            raise JavaStreamError("proxy classes are not part of this format")
        if tc != TC_CLASSDESC:
            raise JavaStreamError(f"expected a class descriptor, got 0x{tc:02x}")
        name, uid = self._utf(), self._u(">q")
        cd = ClassDesc(name, uid, 0, [], None)
        self._new_handle(cd)
        cd.flags = self._u(">B")
        for _ in range(self._u(">H")):
            t = chr(self._u(">B"))
            fname = self._utf()
            cname = self.read_content() if t in "L[" else None
            cd.fields.append((t, fname, cname))
        self._skip_annotation()
        cd.super = self.read_class_desc()
        return cd

    def _ref(self):
        h = self._u(">I") - BASE_HANDLE
        if not 0 <= h < len(self.handles):
            raise JavaStreamError("bad back-reference")
        return self.handles[h]

    def _skip_annotation(self):
        out = []
        while True:
            if self.b[self.p] == TC_ENDBLOCKDATA:
                self.p += 1
                return out
            out.append(self.read_content())

    def read_content(self):
        tc = self._u(">B")
        if tc == TC_NULL:
            return None
        if tc == TC_REFERENCE:
            return self._ref()
        if tc == TC_STRING or tc == TC_LONGSTRING:
            s = self._utf(long=tc == TC_LONGSTRING)
            self._new_handle(s)
            return s
        if tc == TC_BLOCKDATA:
            return self._take(self._u(">B"))
        if tc == TC_BLOCKDATALONG:
            return self._take(self._u(">I"))
        if tc == TC_CLASSDESC or tc == TC_PROXYCLASSDESC:
            self.p -= 1
            return self.read_class_desc()
        if tc == TC_CLASS:
            cd = self.read_class_desc()
            self._new_handle(cd)
            return cd
        if tc == TC_ARRAY:
            cd = self.read_class_desc()
            slot = self._new_handle(None)
            n = self._u(">i")
            et = cd.name[1]
            if et == "B":
                arr = bytes(self._take(n))
            elif et in _PRIM:
                arr = [self._u(_PRIM[et]) for _ in range(n)]
            else:
                arr = [self.read_content() for _ in range(n)]
            self.handles[slot] = arr
            return arr
        if tc == TC_ENUM:
            cd = self.read_class_desc()
            slot = self._new_handle(None)
            self.handles[slot] = (cd.name, self.read_content())
            return self.handles[slot]
        if tc == TC_OBJECT:
            cd = self.read_class_desc()
            obj = JavaObject(cd, {}, [])
            self._new_handle(obj)
            chain = []
            c = cd
            while c is not None:
                chain.append(c)
                c = c.super
            for c in reversed(chain):                               # class data is written from the top-most serializable superclass down
                if c.flags & SC_SERIALIZABLE:
                    for t, fname, _ in c.fields:
                        obj.fields[fname] = self._u(_PRIM[t]) if t in _PRIM else self.read_content()
                    if c.flags & SC_WRITE_METHOD:
                        obj.annotations.extend(self._skip_annotation())
                elif c.flags & SC_EXTERNALIZABLE:
                    if not c.flags & SC_BLOCK_DATA:
                        raise JavaStreamError("protocol-1 externalizable data cannot be skipped")
                    obj.annotations.extend(self._skip_annotation())
            return obj
        raise JavaStreamError(f"unsupported type code 0x{tc:02x}")


@dataclass
class EncryptedPointRecord:
    """The fields of com.fspann.common.EncryptedPoint (EP:18-26) the hot path needs."""
    id: str
    version: int
    iv: bytes
    ciphertext: bytes
    key_version: int
    dimension: int
    shard_id: int


def parse_encrypted_point(data: bytes) -> EncryptedPointRecord:
    obj = JavaObjectStream(data).read_content()
    if not isinstance(obj, JavaObject) or obj.cls.name != "com.fspann.common.EncryptedPoint":
        raise JavaStreamError("stream does not hold a com.fspann.common.EncryptedPoint")      # PersistenceUtils.loadObject's type check (:78-81)
    f = obj.fields
    for k in ("id", "iv", "ciphertext"):
        if f.get(k) is None:
            raise JavaStreamError(f"{k} cannot be null")                                         # EP:41-44
    rec = EncryptedPointRecord(f["id"], int(f["version"]), bytes(f["iv"]), bytes(f["ciphertext"]), int(f["keyVersion"]), int(f["dimension"]),
                               int(f.get("shardId", 0)))
    if len(rec.iv) != 12 or len(rec.ciphertext) != 8 * rec.dimension + 16:
        raise JavaStreamError(f"record {rec.id}: iv/ciphertext lengths {len(rec.iv)}/{len(rec.ciphertext)} do not match dimension {rec.dimension}")
    return rec


def load_point_file(path: str) -> EncryptedPointRecord:
    with open(path, "rb") as fh:
        return parse_encrypted_point(fh.read())


def parse_vector_metadata(value: str) -> dict:
    """RocksDB value of a vector id (RDB:743-752): 'k=v;k=v' where '=' and ';' inside keys / values are backslash-escaped."""
    out = {}
    for tok in re.split(r"(?<!\\);", value):
        kv = re.split(r"(?<!\\)=", tok, maxsplit=1)
        if len(kv) == 2:
            out[kv[0].replace("\\=", "=").replace("\\;", ";")] = kv[1].replace("\\=", "=").replace("\\;", ";")
    return out


def scan_points_dir(base_dir: str, metadata: dict | None = None):
    """Collects the current record of every id under <base_dir>/v<version>/<id>.point.

    metadata: optional {id: 'version=..;shardId=..;dim=..'} (RocksDB values); when given, the file is resolved exactly like
    loadEncryptedPoint (RDB:530-544: version from the metadata, missing file -> id skipped); otherwise the highest version directory
    holding the id wins.  Ids must be the decimal ordinals 0..N-1 the facade assigns (FSA:501,515).
    Returns (iv uint8 [N,12], ct uint8 [N,8*dim+16], key_version int32 [N], present bool [N], dim)."""
    by_id = {}
    for entry in sorted(os.listdir(base_dir)):
        m = re.fullmatch(r"v(-?\d+)", entry)
        d = os.path.join(base_dir, entry)
        if not m or not os.path.isdir(d):
            continue
        ver = int(m.group(1))
        for fn in os.listdir(d):
            if fn.endswith(".point"):
                by_id.setdefault(fn[:-6], {})[ver] = os.path.join(d, fn)
    chosen = {}
    for sid, files in by_id.items():
        if not re.fullmatch(r"\d+", sid):
            continue
        if metadata is not None:
            meta = parse_vector_metadata(metadata.get(sid, ""))
            if "version" not in meta:
                continue                                             # RDB:533-535
            v = meta["version"]
            v = int(v[1:] if v.startswith("v") else v)
            if v not in files:
                continue                                             # RDB:540-542
            chosen[int(sid)] = files[v]
        else:
            chosen[int(sid)] = files[max(files)]
    if not chosen:
        return np.zeros((0, 12), np.uint8), np.zeros((0, 16), np.uint8), np.zeros(0, np.int32), np.zeros(0, bool), 0
    n = max(chosen) + 1
    first = load_point_file(next(iter(chosen.values())))
    dim = first.dimension
    iv = np.zeros((n, 12), dtype=np.uint8)
    ct = np.zeros((n, 8 * dim + 16), dtype=np.uint8)
    kv = np.zeros(n, dtype=np.int32)
    present = np.zeros(n, dtype=bool)
    for i, path in chosen.items():
        try:
            rec = load_point_file(path)
        except (OSError, JavaStreamError):
            continue                                                 # loadPointIfActive swallows load failures (PIS:717-724)
        if rec.dimension != dim or rec.id != str(i):
            continue
        iv[i] = np.frombuffer(rec.iv, dtype=np.uint8)
        ct[i] = np.frombuffer(rec.ciphertext, dtype=np.uint8)
        kv[i] = rec.key_version
        present[i] = True
    return iv, ct, kv, present, dim
