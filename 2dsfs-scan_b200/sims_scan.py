"""Drop-in replacement for the function API of scripts/sims_scan.py (uricchio/2DSFS-scan), backed by CUDA kernels for B200.

Same function names, parameter order and return shapes as the reference's free functions (reference lines 18-690).
The twins differ from the class on purpose, and so do these: no None guards (an empty window raises
ZeroDivisionError, :348/:415), T2D_diff with a MINUS sign (:497), window_end = start + W (:504), the 1D backgrounds
of likelihood_scan left unfolded (:616-617).  All numerics run on the GPU through libtdsfs.so; no CPU fallback.
"""
from __future__ import annotations

import csv
import glob
import os

import numpy as np

import tdsfs_capi as T
from tdsfs_engine import (Engine, SnpTable, dense2d_to_dict, dict_to_dense2d, dict_to_folded1d, filter_flags, stat_lists,
                          window_keys)
from twoDSFS_class import parse_vcf_to_dict

_engine = None


def _eng():
    global _engine
    if _engine is None:
        _engine = Engine(int(os.environ.get("TDSFS_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _engine


def make_data_dict_vcf(vcf_filename, popinfo_filename):
    """reference :18-120"""
    return parse_vcf_to_dict(vcf_filename, popinfo_filename)


def calculate_2d_sfs(data_dict, pop1, pop2, pop1_size, pop2_size, start_position, end_position, variant_type, fold=True):
    """reference :123-234"""
    table = SnpTable.from_dict(data_dict, pop1, pop2)
    h2, _, _ = _eng().spectra(table, pop1_size, pop2_size, fold, filter_flags(table, start_position, end_position, variant_type))
    return dense2d_to_dict(h2.astype(np.int64))


def normalize_2d_sfs(sfs):
    counts = list(sfs.values())
    total = sum(counts[1:-1])
    return {coords: values / total for coords, values in sfs.items()}


def count_snps(window_data, variant_type):
    return sum(1 for v in window_data.values() if variant_type is None or v.get("annotation") == variant_type)


def calculate_1d_sfs(data_dict, pop, pop_size, start_position, end_position, variant_type):
    """reference :262-302"""
    table = SnpTable.from_dict(data_dict, pop, pop)
    _, s1, _ = _eng().spectra(table, pop_size, pop_size, False, filter_flags(table, start_position, end_position, variant_type))
    return {i: int(v) for i, v in enumerate(s1.tolist())}


def fold_1d_sfs(sfs_dict):
    num_chromosomes = max(sfs_dict.keys())
    folded = {}
    for freq, count in sfs_dict.items():
        m = min(freq, num_chromosomes - freq)
        folded[m] = folded[m] + count if m in folded else count
    return folded


def calculate_likelihood_1D(foreground_sfs, background_sfs):
    """reference :325-395 (no guards: ZeroDivisionError on an empty spectrum)"""
    return _eng().likelihood(foreground_sfs, background_sfs, guarded=False)


def calculate_likelihood_2D(foreground_2d_sfs, background_2d_sfs):
    """reference :398-440"""
    return _eng().likelihood(foreground_2d_sfs, background_2d_sfs, guarded=False)


def get_gens(main_dir):
    search_strings = set()
    for root, dirs, files in os.walk(main_dir):
        for file in files:
            parts = file.split('.')
            if len(parts) == 5:
                search_strings.add(parts[1])
    return search_strings


def _unguarded(T2D, T1, T2_, pop1_size, pop2_size):
    """The twins have no None guards: an empty spectrum divides by zero (:348, :415) -- except a 1D spectrum of a single
    diploid, whose interior is empty: nothing is divided and scipy's logpmf of empty vectors gives NaN."""
    if pop1_size <= 1:
        T1 = float("nan")
    if pop2_size <= 1:
        T2_ = float("nan")
    if T2D is None or T1 is None or T2_ is None:
        raise ZeroDivisionError("division by zero")
    return T2D, T1, T2_


def process_window(data_dict, bg_2d_sfs, bg_p1_sfs, bg_p2_sfs, window_size, pop1, pop2, pop1_size, pop2_size, start_position,
                   end_position, variant_type):
    """reference :451-590: fixed-bp scan of one replicate against precomputed backgrounds."""
    table = SnpTable.from_dict(data_dict, pop1, pop2)
    if table.n == 0:
        return {}
    eng = _eng()
    b2 = dict_to_dense2d(bg_2d_sfs, pop1_size, pop2_size)
    b1a = dict_to_folded1d(bg_p1_sfs, pop1_size)
    b1b = dict_to_folded1d(bg_p2_sfs, pop2_size)
    eng.load(table, pop1_size, pop2_size, True, filter_flags(table, start_position, end_position, variant_type))
    eng.background(T.BG_NONE)
    eng.h.set_background(b2, b1a, b1b)
    res = eng.scan(window_size, False)
    live = (res["flags"] & T.F_EMPTY) == 0
    keys = window_keys(table, res, live)
    counts = res["snp_count"][live].tolist()
    starts = res["start"][live].tolist()
    T2, Ta, Tb = stat_lists(res, live)
    results = {}
    for k, c, st, T2D, T1, T2_ in zip(keys, counts, starts, T2, Ta, Tb):
        T2D, T1, T2_ = _unguarded(T2D, T1, T2_, pop1_size, pop2_size)
        results[k] = {"window_type": "background" if 0 <= st < 500000 else "foreground", "window_start": st,
                      "window_end": st + window_size, "snp_count": c, "T2D": T2D, "T1D_p1": T1, "T1D_p2": T2_,
                      "new_term_p1": T2D - T1, "new_term_p2": T2D - T2_, "T2D_diff": T2D - (T1 - T2_) / 2}
    return results


def process_window_batch(data_dicts, bg_2d_sfs, bg_p1_sfs, bg_p2_sfs, window_size, pop1, pop2, pop1_size, pop2_size, start_position,
                         end_position, variant_type):
    """All replicates of one generation in ONE launch (BASELINE.json configs[2]): every replicate becomes a block of
    pseudo-chromosomes, so the 2-3 windows x hundreds of replicates are scored by one pass of the count / boundary / score
    kernels instead of one latency-bound round trip per replicate.  Returns [process_window(d, ...) for d in data_dicts]."""
    tables = [SnpTable.from_dict(d, pop1, pop2) for d in data_dicts]
    big = SnpTable()
    big.chroms, offs, base = [], [0], 0
    for r, t in enumerate(tables):
        big.chroms += [(r, c) for c in t.chroms]
        offs += (t.off[1:] + base).tolist()
        base += t.n
    big.off = np.asarray(offs, dtype=np.int64)
    big.pos = np.concatenate([t.pos for t in tables]) if tables else np.zeros(0, np.int64)
    big.cnt = np.concatenate([t.cnt for t in tables]) if tables else np.zeros((0, 4), np.int64)
    big.ann = np.concatenate([t.ann for t in tables]) if tables else np.array([], dtype=object)
    big.keys, big.pops, big.n, big.last_key_row = [], (pop1, pop2), base, -1
    out = [{} for _ in tables]
    if base == 0:
        return out
    eng = _eng()
    eng.load(big, pop1_size, pop2_size, True, filter_flags(big, start_position, end_position, variant_type))
    eng.background(T.BG_NONE)
    eng.h.set_background(dict_to_dense2d(bg_2d_sfs, pop1_size, pop2_size), dict_to_folded1d(bg_p1_sfs, pop1_size),
                         dict_to_folded1d(bg_p2_sfs, pop2_size))
    res = eng.scan(window_size, False)
    live = (res["flags"] & T.F_EMPTY) == 0
    T2, Ta, Tb = stat_lists(res, live)
    for ci, c, st, en, T2D, T1, T2_ in zip(res["chrom"][live].tolist(), res["snp_count"][live].tolist(), res["start"][live].tolist(),
                                           res["end"][live].tolist(), T2, Ta, Tb):
        T2D, T1, T2_ = _unguarded(T2D, T1, T2_, pop1_size, pop2_size)
        r, chrom = big.chroms[ci]
        out[r][f"{chrom} {st}-{en}"] = {"window_type": "background" if 0 <= st < 500000 else "foreground", "window_start": st,
                                        "window_end": st + window_size, "snp_count": c, "T2D": T2D, "T1D_p1": T1, "T1D_p2": T2_,
                                        "new_term_p1": T2D - T1, "new_term_p2": T2D - T2_, "T2D_diff": T2D - (T1 - T2_) / 2}
    return out


POPMAP_SIMS = "/Users/marlonalejandrocalderonbalcazar/Desktop/ECB/simulations/results/popmap_sims_copy.txt"


def _iter_replicates(main_dir, popinfo_filename):
    for generation in get_gens(main_dir):
        target_vcfs = glob.glob(f"{main_dir}/iter*/*{generation}*.vcf.gz")
        concatenated_vcfs = glob.glob(f"{main_dir}/concatenated_vcfs/gen.{generation}.concatenated.vcf.gz")
        for vcf in concatenated_vcfs:
            data_dict = make_data_dict_vcf(vcf, popinfo_filename)
            bg_2d_sfs = calculate_2d_sfs(data_dict, 'p1', 'p2', 5, 5, start_position=0, end_position=500000, variant_type=None)
            bg_p1_sfs = calculate_1d_sfs(data_dict, 'p1', 5, start_position=0, end_position=500000, variant_type=None)
            bg_p2_sfs = calculate_1d_sfs(data_dict, 'p2', 5, start_position=0, end_position=500000, variant_type=None)
            dicts = [make_data_dict_vcf(v, popinfo_filename) for v in target_vcfs]
            batch = process_window_batch(dicts, bg_2d_sfs, bg_p1_sfs, bg_p2_sfs, 500000, 'p1', 'p2', 5, 5,
                                         start_position=None, end_position=None, variant_type=None)
            for vcf_input, results in zip(target_vcfs, batch):
                yield generation, int(vcf_input.split('.')[2]), results


def likelihood_scan(main_dir, popinfo_filename=None):
    """reference :646-690 (the definition that is live at import; it shadows the CSV-writing one at :593).
    popinfo_filename: the reference hard-codes a path on the author's machine; pass yours here."""
    popinfo_filename = popinfo_filename or POPMAP_SIMS
    likelihood_results = {}
    for generation, iteration_number, results in _iter_replicates(main_dir, popinfo_filename):
        for key, value in results.items():
            window_start, window_end = map(int, key.split(' ')[1].split('-'))
            region = 'background' if window_end <= 1000000 else 'foreground'
            likelihood_results[(generation, iteration_number, key)] = {'generation': generation, 'iteration': iteration_number,
                                                                        'region': region, 'window_coords': key, 'likelihood': value}
    return likelihood_results


def likelihood_scan_to_csv(main_dir, output, popinfo_filename=None):
    """The shadowed first definition of likelihood_scan (reference :593-644): one CSV row per window."""
    popinfo_filename = popinfo_filename or POPMAP_SIMS
    cols = ['generation', 'iteration', 'region', 'window_coords', 'snp_count', 'T2D', 'T1D_p1', 'T1D_p2', 'new_term_p1',
            'new_term_p2', 'T2D_diff']
    with open(output, 'w', newline='') as csvfile:
        writer = csv.DictWriter(csvfile, fieldnames=cols)
        writer.writeheader()
        for generation, iteration_number, results in _iter_replicates(main_dir, popinfo_filename):
            for window_coords, result in results.items():
                window_start, window_end = window_coords.split(' ')[1].split('-')
                region = 'background' if int(window_end) <= 1000000 else 'foreground'
                writer.writerow({'generation': generation, 'iteration': iteration_number, 'region': region,
                                 'window_coords': window_coords, 'snp_count': result["snp_count"], 'T2D': result["T2D"],
                                 'T1D_p1': result["T1D_p1"], 'T1D_p2': result["T1D_p2"], 'new_term_p1': result["new_term_p1"],
                                 'new_term_p2': result["new_term_p2"], 'T2D_diff': result["T2D_diff"]})
