#!/usr/bin/env python
"""SASS evidence for profiles/: Blackwell-native / asynchronous-copy instructions per kernel of libfspann_gpu.so (cuobjdump -sass).

  python tools/sass_evidence.py > profiles/r2_sass_blackwell.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fspann_query_system_b200", "csrc", "libfspann_gpu.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCIMMA|LDTM[.\w]*|STTM[.\w]*|UTCBAR|UTCATOMSWS[.\w]*|UBLKCP[.\w]*|UTMALDG[.\w]*|SYNCS[.\w]*|LDGSTS[.\w]*|LDGDEPBAR|DEPBAR[.\w]*|REDUX[.\w]*|ATOMS[.\w]*|VIMNMX[.\w]*|VABSDIFF4[.\w]*|IDP[.\w]*)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    fn, rows = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m or fn is None:
            continue
        ins = m.group(1).strip()
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        mm = PAT.match(ins)
        if mm:
            key = (fn, mm.group(1))
            cnt, ex = rows.get(key, (0, ins))
            rows[key] = (cnt + 1, ex)
    print("# r2 SASS evidence: cuobjdump -sass fspann_query_system_b200/csrc/libfspann_gpu.so (sm_100a); tools/sass_evidence.py")
    print("# UTCHMMA = tcgen05.mma (kind::f16, BF16 operands) | LDTM = tcgen05.ld (TMEM -> registers) | UTCBAR = tcgen05.commit -> mbarrier |")
    print("# UTCATOMSWS = tcgen05.alloc / relinquish / dealloc | UBLKCP.S.G = cp.async.bulk global -> shared (TMA bulk copy) | SYNCS.* = mbarrier ops |")
    print("# LDGSTS = cp.async (global -> shared without registers) | ATOMS = shared-memory atomics (POPC.INC / OR / CAS) | VIMNMX = branch-free sort compare |")
    print("# VABSDIFF4 / IDP = byte distance path (|q - v| on 4 bytes, dot product accumulate)")
    print(f"{'kernel':52s} {'mnemonic':34s} {'count':>5s}  example")
    for (fn, mn), (cnt, ex) in rows.items():
        if fn.startswith("void "):
            fn = fn[5:]
        if not fn.startswith("fsp::") or "cub::" in fn:
            continue
        print(f"{fn[:52]:52s} {mn:34s} {cnt:5d}  {ex[:90]}")


if __name__ == "__main__":
    main()
