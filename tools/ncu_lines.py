#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py <rep> <kernel-regex> [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = {}
for r in rows:
    if len(r) > ie and r[0].isdigit() and r[2] == "-":  # per-line summary rows
        k = (int(r[0]), r[1].strip())
        a = agg.setdefault(k, [0, 0])
        a[0] += int(r[ie] or 0); a[1] += int(r[isamp] or 0)
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instructions {tot}, samples {tots}")
for (ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} {n:10d} {100*n/tot:5.1f}%  samp {100*s/tots:5.1f}%  {src[:105]}")
