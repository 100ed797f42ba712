#!/usr/bin/env python
"""Top CUDA source lines of one kernel in an .ncu-rep (stall samples / instructions executed).
usage: python tools/ncu_lines.py rep.ncu-rep kernel_regex [topN]"""
import csv, io, subprocess, sys
from collections import defaultdict

def num(s):
    try: return int(float(s))
    except Exception: return 0

def main():
    rep, k = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + k],
                         capture_output=True, text=True).stdout
    cur, hdr = None, None
    agg = defaultdict(lambda: [0, 0, ""])
    for r in csv.reader(io.StringIO(txt)):
        if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if len(r) > 5 and r[0] == "Line No": hdr = r; continue
        if hdr and len(r) > 8 and r[0].isdigit():
            a = agg[(cur, int(r[0]))]
            a[0] += num(r[hdr.index("# Samples")]); a[1] += num(r[hdr.index("Instructions Executed")]); a[2] = r[1].strip()[:120]
    ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
    print("samples", ts, "instructions", ti)
    for (f, l), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print("%-18s %4d  smp %5.1f%%  inst %5.1f%%  %s" % (f, l, 100 * s / ts, 100 * i / ti, src))

if __name__ == "__main__":
    main()
