"""Drop-in for the free functions of the first-generation script scripts/twoDSFS.py of uricchio/2DSFS-scan that form its
Poisson composite scan: calculate_2d_sfs (:211-303), normalize_2d_sfs (:324-334),
calculate_p (:336-374), count_snps (:376-383) and calculate_p_window (:385-463).  Same names, parameter order and return
shapes.  The class copy of calculate_p_window (scripts/src/twoDSFS_class.py:304-393) cannot run (it calls
calculate_2d_sfs with nine arguments); this script is the version that produced results.

Every spectrum, window assignment and score runs on the GPU through libtdsfs.so (tdsfs_set_poisson_background +
tdsfs_scan_poisson_bp, include/tdsfs.h); this file adapts dicts to arrays and back.  No CPU fallback."""
from __future__ import annotations

import os

import numpy as np

import tdsfs_capi as T
from tdsfs_engine import Engine, SnpTable, filter_flags, window_keys

_engine = None


def _eng():
    global _engine
    if _engine is None:
        _engine = Engine(int(os.environ.get("TDSFS_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _engine


def calculate_2d_sfs(data_dict, pop1, pop2, pop1_size, pop2_size, start_position, end_position, variant_type=None):
    """:211-303.  UNFOLDED joint spectrum keyed by raw alt counts (SNPs with both alt counts 0 skipped), dense dict in row-major
    insertion order, plus a pseudo-count 1 / total_sites on every bin (0 when no site was counted)."""
    table = SnpTable.from_dict(data_dict, pop1, pop2)
    flags = filter_flags(table, start_position, end_position, variant_type)
    h2, _, _ = _eng().spectra(table, pop1_size, pop2_size, False, flags)
    h2 = h2.astype(np.int64)
    R1, R2 = h2.shape
    total = int(h2.sum())
    pc = 1 / total if total > 0 else 0
    flat = h2.ravel().tolist()
    return {(i, j): flat[i * R2 + j] + pc for i in range(R1) for j in range(R2)}


def normalize_2d_sfs(sfs):
    """:324-334: divide every bin by the sum of all bins but the first and the last (in insertion order)."""
    counts = list(sfs.values())
    total = sum(counts[1:-1])
    return {coords: values / total for coords, values in sfs.items()}


def calculate_p(foreground_sfs, background_sfs):
    """:336-374: sum over the bins with a non-zero expectation S_w * bg[k] of poisson.logpmf(int(fg[k]), S_w * bg[k])."""
    keys = list(foreground_sfs.keys())
    S_w = sum(foreground_sfs.values())
    x = np.array([int(foreground_sfs[k]) for k in keys], dtype=np.int64)
    mu = np.array([S_w * background_sfs.get(k, 0) for k in keys], dtype=np.float64)
    return _eng().h.poisson_score(x, mu)


def count_snps(window_data, variant_type):
    """:376-383."""
    return sum(1 for v in window_data.values() if variant_type is None or v.get("annotation") == variant_type)


def calculate_p_window(data_dict, sfs_normalized, window_size, pop1, pop2, pop1_size, pop2_size, start_position, end_position,
                       variant_type):
    """:385-463.  {'<chrom> <start>-<end>': {'p_values': score, 'snp_count': n}} for every non-empty fixed-bp window, in the
    sorted (chromosome string, position) order of the reference's walk."""
    eng = _eng()
    table = SnpTable.from_dict(data_dict, pop1, pop2)
    if table.n == 0:
        return {}
    flags = filter_flags(table, start_position, end_position, variant_type)
    eng.load(table, pop1_size, pop2_size, False, flags)
    eng.background(T.BG_NONE)
    R1, R2 = 2 * pop1_size + 1, 2 * pop2_size + 1
    q = np.zeros(R1 * R2, dtype=np.float64)
    for (i, j), v in sfs_normalized.items():      # background_sfs.get(k, 0) for the keys of the dense foreground (:347)
        if 0 <= i < R1 and 0 <= j < R2:
            q[i * R2 + j] = v
    eng.h.set_poisson_background(q)
    res = eng.h.scan_poisson(window_size)
    live = (res["flags"] & T.F_EMPTY) == 0
    keys = window_keys(table, res, live)
    return {k: {"p_values": p, "snp_count": c} for k, p, c in zip(keys, res["T2D"][live].tolist(), res["snp_count"][live].tolist())}
