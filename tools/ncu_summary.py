#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key metrics per kernel + opcode mix + top stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof_c5.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
seen = set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if name in seen:
        continue
    seen.add(name)
    print("====", name)
    for w in want:
        if w in idx:
            print(f"  {w} = {r[idx[w]]} {rows[1][idx[w]]}")
    vals = [(float(r[idx[h]] or 0), h[33:]) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    tot = sum(v for v, _ in vals) or 1
    print("  stalls:", ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(vals, reverse=True)[:7]))
pat = sys.argv[2] if len(sys.argv) > 2 else None
if pat:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    body = []
    for r in rows[2:]:
        if len(r) < 10 or r[0] == "Kernel Name":
            break
        body.append(r)
    tot = sum(int(r[ie]) for r in body)
    c, s = Counter(), Counter()
    for r in body:
        t = r[ia].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[ie]); s[op] += int(r[isamp])
    print(f"  SASS instrs {len(body)}, executed {tot}")
    for op, v in c.most_common(18):
        print(f"    {op:8s} {v:11d} {100 * v / tot:5.1f}%  samples {s[op]}")
    top = sorted(body, key=lambda r: -int(r[isamp]))[:25]
    print("  hottest SASS by samples:")
    for r in top:
        print(f"    {r[isamp]:>6} {r[ie]:>9}  {r[ia].strip()[:90]}")
