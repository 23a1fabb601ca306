#!/usr/bin/env python
"""Latency of the reference-sized cases (BASELINE.json configs[0..2]) through the drop-in Python API on one GPU:
chr1 golden dict (418,367 SNPs, 18+14 diploids) at 20 kb / 500 kb / 500-SNP windows, and a batched sims generation.
Prints one JSON line per case: wall seconds split into dict->array conversion, GPU calls, result dict building."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "2dsfs-scan_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import twoDSFS_class as K  # noqa: E402
import tdsfs_capi as T  # noqa: E402
from helpers import load_chr1_arrays, load_chr1_dict  # noqa: E402

d = load_chr1_dict()
inst = K.LikelihoodInference_jointSFS("x", "y")
inst.combined_scan(d, 20000)  # warm-up (library load, context, first-touch)
for name, fn, arg in (("combined_scan 20 kb", inst.combined_scan, 20000), ("combined_scan 500 kb", inst.combined_scan, 500000),
                      ("scan_perChr_bySNPs 500", inst.scan_perChr_bySNPs, 500)):
    inst._tcache = None
    t0 = time.perf_counter()
    tab = inst._table(d)
    t1 = time.perf_counter()
    res = fn(d, arg)
    t2 = time.perf_counter()
    print(json.dumps({"case": f"chr1 golden dict, {name}", "snps": len(d), "windows": len(res), "dict_to_arrays_s": round(t1 - t0, 4),
                      "scan_call_s": round(t2 - t1, 4), "snps_per_s_api": round(len(d) / (t2 - t0)), "reference_cpu_s": "5.2 (BASELINE.md, 1 core)"}))
# C-ABI only (arrays already built): counts entry
chrom, pos, cnt, ann, vocab = load_chr1_arrays()
h = T.Handle(0)
h.set_panel(18, 14, True)
h.load_counts(cnt, pos, [0, len(pos)])
for W in (20000, 500000):
    h.run_bp(T.BG_PER_CHROM, W)
    t0 = time.perf_counter()
    for _ in range(20):
        h.load_counts(cnt, pos, [0, len(pos)])
        r = h.run_bp(T.BG_PER_CHROM, W)
    dt = (time.perf_counter() - t0) / 20
    print(json.dumps({"case": f"chr1 arrays through the C ABI (host buffers in, host results out), {W} bp", "snps": len(pos), "seconds": round(dt, 6),
                      "snps_per_s": round(len(pos) / dt), "kernel_ms": {k: round(v, 4) for k, v in h.timings().items()}}))
