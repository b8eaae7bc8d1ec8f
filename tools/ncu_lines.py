"""Per-source-line share of executed instructions and stall samples from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [min_pct] [kernel-regex]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if len(sys.argv) > 3:
        cmd += ["--kernel-name", "regex:" + sys.argv[3]]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = [r for r in rows if r and r[0] == "Line No"][0]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    sec, agg, tot, tots = None, collections.OrderedDict(), 0, 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            sec = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No") or r[0] == "" or sec is None:
            continue
        try:
            n, smp = int(r[ci]), int(r[cs])
        except ValueError:
            continue
        agg[(sec, int(r[0]))] = (n, smp, r[1])
        tot += n
        tots += smp
    print("total warp-instructions", tot, "samples", tots)
    for (f, ln), (n, smp, src) in sorted(agg.items()):
        if n > tot * thr / 100 or smp > tots * thr / 100:
            print(f"{f}:{ln:4d} {n / tot * 100:5.1f}% inst {smp / tots * 100:5.1f}% smp  {src.strip()[:100]}")


if __name__ == "__main__":
    main()
