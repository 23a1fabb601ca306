#!/usr/bin/env python
"""SASS opcode counts per kernel of the shipped library (no GPU needed): python tools/sass_summary.py > profiles/rNN_sass_summary.txt
UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier operations, ATOMS / ATOMG / RED / REDG = shared / global atomics,
LDG/STG with peer pointers are the NVLink traffic of the exchange kernels."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "2dsfs-scan_b200", "lib", "libtdsfs.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = {}
try:
    names = sorted(set(re.findall(r"Function : (\S+)", out)))
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    demangle = dict(zip(names, dm))
except Exception:  # noqa: BLE001
    pass
cols = ["UBLKCP", "SYNCS", "LOP3", "POPC", "ATOMS", "ATOMG", "RED", "REDG", "DFMA", "DADD", "LDS", "LDG", "STG", "BRA"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        counts[cur]["_all"] += 1
print(f"SASS opcode counts per kernel (cuobjdump -sass {os.path.relpath(lib, ROOT)}; sm_100a)")
print("UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops")
print()
print(f"{'kernel':78s} {'instrs':>6s} " + " ".join(f"{c:>6s}" for c in cols))
for k, c in counts.items():
    name = demangle.get(k, k)
    name = re.sub(r"\(.*\)$", "", name).replace("void ", "").replace("tdsfs::", "")
    print(f"{name[:78]:78s} {c['_all']:6d} " + " ".join(f"{c[x]:6d}" for x in cols))
