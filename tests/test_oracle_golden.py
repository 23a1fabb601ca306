"""Pin the CPU oracle (oracle/sfs_oracle.py) against the reference's own golden vectors.

(1) shipped outputs data/ECBstats_*.csv (chromosome 1) from data/chr1.pkl.bz2;
(2) the unmodified reference executed in the build container on seeded inputs
    (tests/golden/make_golden.py), including the exceptions it raises.
CPU only."""
import json
import os

import numpy as np
import pytest

import sfs_oracle as O
from helpers import GOLDEN, close, compare_result_lists, load_chr1_arrays, load_chr1_dict, load_ecb_csv, load_small, rows_to_dict


@pytest.fixture(scope="module")
def chr1():
    return load_chr1_dict()


def _check_csv(res, rows, chrom):
    # data/ECBstats_{20kb,500kb}.csv were re-written (and re-ordered) by scripts/ECBstats_plots.R:239-240 -> match by window
    assert len(res) == len(rows)
    assert set(res) == {f"{chrom} {r['window_start']}-{r['window_end']}" for r in rows}
    for r in rows:
        k = f"{chrom} {r['window_start']}-{r['window_end']}"
        v = res[k]
        assert v["snp_count"] == r["snp_count"]
        for a, b in (("T2D", "T2D"), ("T1D_pop1", "T1D_p1"), ("T1D_pop2", "T1D_p2"), ("new_term_pop1", "new_term_p1"),
                     ("new_term_pop2", "new_term_p2"), ("T2D_diff", "T2D_diff")):
            assert close(v[a], r[b]), (k, a, v[a], r[b])


@pytest.mark.parametrize("tag,W", [("20kb", 20000), ("500kb", 500000)])
def test_shipped_csv_fixed_bp(chr1, tag, W):
    res = O.combined_scan(O.Panel(), chr1, W)
    _check_csv(res, load_ecb_csv(tag), "NC_087088.1")


def test_shipped_csv_fixed_snp(chr1):
    res = O.scan_perChr_bySNPs(O.Panel(), chr1, 500)
    _check_csv(res, load_ecb_csv("500snps"), "NC_087088.1")


def test_stale_carry_row_present():
    rows = load_ecb_csv("20kb")
    r = [x for x in rows if x["window_start"] == 40001][0]
    prev = [x for x in rows if x["window_start"] == 20001][0]
    assert r["T2D"] is None and r["new_term_p1"] == prev["new_term_p1"]  # SURVEY 9.Q6 is in the golden data


def _mk_panel(ctor):
    return O.Panel(**ctor)


def _run(case, call):
    p = _mk_panel(case["ctor"])
    d = rows_to_dict(case["rows"], case["pops"])
    m, a = call["method"], call["args"]
    raw2 = O.calculate_2d_sfs(d, p.pop1, p.pop2, p.n1, p.n2, p.start, p.end, p.vt, p.fold)
    f1 = lambda: p.s1(d, 1)  # noqa: E731
    f2 = lambda: p.s1(d, 2)  # noqa: E731
    if m == "calculate_2d_sfs":
        return raw2
    if m == "calculate_1d_sfs":
        return O.calculate_1d_sfs(d, a[0], a[1], p.start, p.end, p.vt)
    if m == "fold_1d_sfs":
        n = p.n1 if a[0] == p.pop1 else p.n2
        return O.fold_1d_sfs(O.calculate_1d_sfs(d, a[0], n, p.start, p.end, p.vt))
    if m == "normalize_2d_sfs":
        return O.normalize_sfs(raw2)
    if m == "count_snps":
        return O.count_snps(d, a[0])
    if m == "combined_scan":
        return O.combined_scan(p, d, a[0])
    if m == "scan_chooseChr":
        return O.scan_chooseChr(p, d, a[0], a[1])
    if m == "scan_precomputed_BG_raw":
        return O.scan_precomputed_BG(p, d, a[0], raw2, f1(), f2())
    if m == "scan_precomputed_BG_norm":
        return O.scan_precomputed_BG(p, d, a[0], O.normalize_sfs(raw2), O.normalize_sfs(f1()), O.normalize_sfs(f2()))
    if m == "T2D_scan":
        return O.T2D_scan(p, d, raw2, a[0])
    if m == "T1D_scan":
        return O.T1D_scan(p, d, f1(), a[0], a[1], a[2])
    if m == "sims.process_window":
        W, half = a
        s2 = O.calculate_2d_sfs(d, p.pop1, p.pop2, p.n1, p.n2, 0, half, None)
        s1 = O.calculate_1d_sfs(d, p.pop1, p.n1, 0, half, None)
        s1b = O.calculate_1d_sfs(d, p.pop2, p.n2, 0, half, None)
        return O.sims_process_window(d, s2, s1, s1b, W, p.pop1, p.pop2, p.n1, p.n2, None, None, None)
    if m == "sims_process_window":
        return O.sims_process_window_class(p, d, a[0], raw2, f1(), f2())
    if m == "scan_perChr_bySNPs":
        return O.scan_perChr_bySNPs(p, d, a[0])
    if m == "scan_chooseChr_bySNPs":
        return O.scan_chooseChr_bySNPs(p, d, a[0], a[1])
    raise AssertionError(m)


SMALL = load_small()
IDS = [(ci, ki) for ci, c in enumerate(SMALL["cases"]) for ki, _ in enumerate(c["calls"])]


@pytest.mark.parametrize("ci,ki", IDS, ids=[f"{SMALL['cases'][c]['name']}-{SMALL['cases'][c]['calls'][k]['method']}-{k}" for c, k in IDS])
def test_small_cases(ci, ki):
    case = SMALL["cases"][ci]
    call = case["calls"][ki]
    if call["status"] == "raises":
        with pytest.raises(Exception) as ei:
            _run(case, call)
        assert type(ei.value).__name__ == call["result"], (type(ei.value).__name__, call["result"])
        return
    got = _run(case, call)
    exp = call["result"]
    m = call["method"]
    if m in ("calculate_2d_sfs", "normalize_2d_sfs"):
        nz = [[i, j, v] for (i, j), v in got.items() if v != 0]
        assert len(got) == exp[-1][1]
        assert len(nz) == len(exp) - 1
        for (i, j, v), (ei_, ej, ev) in zip(nz, exp[:-1]):
            assert (i, j) == (ei_, ej) and close(v, ev, 1e-15)
    elif m in ("calculate_1d_sfs", "fold_1d_sfs"):
        assert [[k, v] for k, v in got.items()] == exp
    elif m == "count_snps":
        assert got == exp
    else:
        compare_result_lists(got, exp, m)


def test_likelihood_kats():
    for lk in SMALL["likelihood"]:
        fg = dict(enumerate(lk["fg"]))
        bg = dict(enumerate(lk["bg"]))
        assert lk["cls_status"] == "ok"
        assert close(O.likelihood(fg, bg, True), lk["cls"]), lk
        if lk["sims_status"] == "ok":
            assert close(O.likelihood(fg, bg, False), lk["sims"]), lk
        else:
            with pytest.raises(ZeroDivisionError):
                O.likelihood(fg, bg, False)


def test_multinomial_restatement_vs_scipy():
    scipy_stats = pytest.importorskip("scipy.stats")
    rng = np.random.default_rng(3)
    for _ in range(50):
        k = int(rng.integers(1, 40))
        x = rng.poisson(rng.choice([0.2, 5, 100]), k)
        n = int(x.sum())
        p = rng.random(k)
        p /= p.sum()
        if rng.random() < 0.2:
            p[int(rng.integers(k))] = 0.0
            p /= p.sum() if p.sum() > 0 else 1
        a, b = O.multinomial_logpmf(x, n, p), float(scipy_stats.multinomial.logpmf(x, n, p))
        assert close(a, b, 1e-12) or (np.isnan(a) and np.isnan(b)), (a, b)


def test_ingest_small():
    exp = json.load(open(os.path.join(GOLDEN, "ingest_small.json")))["data_dict"]
    d = O.make_data_dict_vcf(os.path.join(GOLDEN, "ingest_small.vcf.gz"), os.path.join(GOLDEN, "ingest_small.popmap.txt"))
    got = [[k, list(v["segregating"]), v["context"], {p: list(c) for p, c in v["calls"].items()}, v["annotation"]] for k, v in d.items()]
    assert got == exp


def test_ingest_ecb_subset_and_scans():
    exp = json.load(open(os.path.join(GOLDEN, "ecb_subset.json")))
    d = O.make_data_dict_vcf(os.path.join(GOLDEN, "ecb_subset.vcf.gz"), os.path.join(GOLDEN, "ecb_subset.popmap.txt"))
    assert len(d) == exp["n_snps"]
    assert [[k, list(v["calls"]["uv"]), list(v["calls"]["bv"])] for k, v in d.items()] == exp["counts"]
    compare_result_lists(O.combined_scan(O.Panel(), d, 20000), exp["combined_20kb"]["result"], "20kb")
    compare_result_lists(O.combined_scan(O.Panel(), d, 500000), exp["combined_500kb"]["result"], "500kb")
    compare_result_lists(O.scan_perChr_bySNPs(O.Panel(), d, 500), exp["bysnps_500"]["result"], "500snps")


def test_array_level_restatement_matches_dict_level():
    """The vectorised array-level scan (used to check the GPU at scale) equals the dict-level restatement."""
    chrom, pos, cnt, ann, vocab = load_chr1_arrays()
    runs = np.load(os.path.join(GOLDEN, "chr1_ref_runs.npz"))
    h2, h1, h1b = O.dense_spectra(cnt, 18, 14)
    assert np.array_equal(h2, runs["bg2d"]) and np.array_equal(h1, runs["bg1d_uv"]) and np.array_equal(h1b, runs["bg1d_bv"])
    off = np.array([0, len(pos)])
    for tag, kw in (("c20k", dict(W=20000)), ("c500k", dict(W=500000)), ("s500", dict(N=500))):
        r = O.scan_arrays(cnt, pos, off, 18, 14, **kw)
        assert np.array_equal(r["start"], runs[f"{tag}_start"]) and np.array_equal(r["end"], runs[f"{tag}_end"])
        assert np.array_equal(r["snp_count"], runs[f"{tag}_snp_count"])
        for a, b in (("T2D", "T2D"), ("T1D_p1", "T1D_pop1"), ("T1D_p2", "T1D_pop2")):
            none = runs[f"{tag}_{b}_none"]
            assert np.array_equal(r[a + "_none"], none)
            ok = ~none
            assert np.all(np.abs(r[a][ok] - runs[f"{tag}_{b}"][ok]) <= 1e-9 * np.maximum(np.abs(runs[f"{tag}_{b}"][ok]), 1))
