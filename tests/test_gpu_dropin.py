"""The reference-facing Python API (2dsfs-scan_b200/twoDSFS_class.py, sims_scan.py) against outputs of the UNMODIFIED
reference recorded in tests/golden/ (make_golden.py): every scanner, spectrum builder and likelihood, including the
exceptions the reference raises, plus the reference's shipped chr1 outputs through the dict API."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, close, compare_result_lists, load_chr1_dict, load_ecb_csv, load_small, rows_to_dict

pytestmark = pytest.mark.gpu

SMALL = load_small()
IDS = [(ci, ki) for ci, c in enumerate(SMALL["cases"]) for ki, _ in enumerate(c["calls"])]


@pytest.fixture(scope="module")
def mods():
    import sims_scan
    import twoDSFS_class
    return twoDSFS_class, sims_scan


def _run(mods, case, call):
    K, S = mods
    ctor = case["ctor"]
    inst = K.LikelihoodInference_jointSFS("x", "y", **ctor)
    d = rows_to_dict(case["rows"], case["pops"])
    m, a = call["method"], call["args"]
    p1, p2, n1, n2 = ctor["pop1"], ctor["pop2"], ctor["pop1_size"], ctor["pop2_size"]
    st, en, vt = ctor.get("start_position"), ctor.get("end_position"), ctor.get("variant_type")

    def fresh():
        return K.LikelihoodInference_jointSFS("x", "y", **ctor)

    def raw2():
        return fresh().calculate_2d_sfs(d)

    def f1(which):
        i = fresh()
        pop, n = (p1, n1) if which == 1 else (p2, n2)
        return i.fold_1d_sfs(i.calculate_1d_sfs(d, pop, n, st, en, vt))

    if m == "calculate_2d_sfs":
        return inst.calculate_2d_sfs(d)
    if m == "calculate_1d_sfs":
        return inst.calculate_1d_sfs(d, a[0], a[1], st, en, vt)
    if m == "fold_1d_sfs":
        n = n1 if a[0] == p1 else n2
        return inst.fold_1d_sfs(inst.calculate_1d_sfs(d, a[0], n, st, en, vt))
    if m == "normalize_2d_sfs":
        return inst.normalize_2d_sfs(raw2())
    if m == "count_snps":
        return inst.count_snps(d, a[0])
    if m == "combined_scan":
        return inst.combined_scan(d, a[0])
    if m == "scan_chooseChr":
        return inst.scan_chooseChr(d, a[0], a[1])
    if m == "scan_precomputed_BG_raw":
        return inst.scan_precomputed_BG(d, a[0], raw2(), f1(1), f1(2))
    if m == "scan_precomputed_BG_norm":
        i = fresh()
        return inst.scan_precomputed_BG(d, a[0], i.normalize_2d_sfs(raw2()), i.normalize_1d_sfs(f1(1)), i.normalize_1d_sfs(f1(2)))
    if m == "T2D_scan":
        return inst.T2D_scan(d, raw2(), a[0])
    if m == "T1D_scan":
        return inst.T1D_scan(d, f1(1), a[0], a[1], a[2])
    if m == "sims.process_window":
        W, half = a
        s2 = S.calculate_2d_sfs(d, p1, p2, n1, n2, 0, half, None)
        s1 = S.calculate_1d_sfs(d, p1, n1, 0, half, None)
        s1b = S.calculate_1d_sfs(d, p2, n2, 0, half, None)
        return S.process_window(d, s2, s1, s1b, W, p1, p2, n1, n2, None, None, None)
    if m == "sims_process_window":
        return inst.sims_process_window(d, a[0], raw2(), f1(1), f1(2))
    if m == "scan_perChr_bySNPs":
        return inst.scan_perChr_bySNPs(d, a[0])
    if m == "scan_chooseChr_bySNPs":
        return inst.scan_chooseChr_bySNPs(d, a[0], a[1])
    raise AssertionError(m)


@pytest.mark.parametrize("ci,ki", IDS, ids=[f"{SMALL['cases'][c]['name']}-{SMALL['cases'][c]['calls'][k]['method']}-{k}" for c, k in IDS])
def test_small_cases(mods, ci, ki):
    case = SMALL["cases"][ci]
    call = case["calls"][ki]
    if call["status"] == "raises":
        with pytest.raises(Exception) as ei:
            _run(mods, case, call)
        assert type(ei.value).__name__ == call["result"], (type(ei.value).__name__, str(ei.value), call["result"])
        return
    got = _run(mods, case, call)
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


def test_likelihood_and_poisson_kats(mods):
    K, S = mods
    inst = K.LikelihoodInference_jointSFS("x", "y")
    for lk in SMALL["likelihood"]:
        fg, bg = dict(enumerate(lk["fg"])), dict(enumerate(lk["bg"]))
        assert close(inst.calculate_likelihood_1D(fg, bg), lk["cls"]), lk
        assert close(inst.calculate_likelihood_2D(fg, bg), lk["cls"]), lk
        if lk["sims_status"] == "ok":
            assert close(S.calculate_likelihood_1D(fg, bg), lk["sims"]), lk
        else:
            with pytest.raises(ZeroDivisionError):
                S.calculate_likelihood_2D(fg, bg)
    for pk in SMALL["poisson"]:
        fg = {(i, j): v for i, j, v in pk["fg"]}
        bg = {(i, j): v for i, j, v in pk["bg"]}
        assert close(inst.calculate_p(fg, bg), pk["value"]), pk
    assert inst.new_term(2.0, 5.5) == 3.5


@pytest.fixture(scope="module")
def chr1():
    return load_chr1_dict()


@pytest.mark.parametrize("tag,method,arg", [("20kb", "combined_scan", 20000), ("500kb", "combined_scan", 500000),
                                            ("500snps", "scan_perChr_bySNPs", 500)])
def test_chr1_shipped_outputs_through_dict_api(mods, chr1, tag, method, arg, tmp_path):
    """data/chr1.pkl.bz2 -> data/ECBstats_*.csv through the drop-in class, stale-carry row included, then the CSV writer."""
    K, _ = mods
    inst = K.LikelihoodInference_jointSFS("x", "y")
    res = getattr(inst, method)(chr1, arg)
    rows = load_ecb_csv(tag)
    chrom = "NC_087088.1"
    assert set(res) == {f"{chrom} {r['window_start']}-{r['window_end']}" for r in rows}
    for r in rows:
        v = res[f"{chrom} {r['window_start']}-{r['window_end']}"]
        assert v["snp_count"] == r["snp_count"]
        for a, b in (("T2D", "T2D"), ("T1D_pop1", "T1D_p1"), ("T1D_pop2", "T1D_p2"), ("new_term_pop1", "new_term_p1"),
                     ("new_term_pop2", "new_term_p2"), ("T2D_diff", "T2D_diff")):
            assert close(v[a], r[b]), (r, a, v[a])
    # CSV writer: same columns / chromosome mapping / '' for None as the reference's save_csv_stats
    K.chr_ids.clear()
    K.chr_ids[chrom] = "1"
    out = tmp_path / "stats.csv"
    K.save_csv_stats(res, str(out))
    lines = out.read_text().splitlines()
    assert lines[0] == "chromosome,window_start,window_end,snp_count,T2D,T1D_p1,T1D_p2,new_term_p1,new_term_p2,T2D_diff"
    assert len(lines) == len(rows) + 1 and lines[1].startswith("1,")
    if tag == "20kb":
        assert any(l.startswith("1,40001,60000,4,,,,") for l in lines)  # the stale-carry row of the golden CSV


def test_ecb_subset_vcf_end_to_end(mods):
    """VCF + popmap -> make_data_dict_vcf -> scans, against the reference run on the same files (BASELINE configs 1-2)."""
    K, _ = mods
    exp = json.load(open(os.path.join(GOLDEN, "ecb_subset.json")))
    inst = K.LikelihoodInference_jointSFS(os.path.join(GOLDEN, "ecb_subset.vcf.gz"), os.path.join(GOLDEN, "ecb_subset.popmap.txt"))
    d = inst.make_data_dict_vcf(inst.vcf_filename, inst.popinfo_filename)
    assert len(d) == exp["n_snps"]
    compare_result_lists(inst.combined_scan(d, 20000), exp["combined_20kb"]["result"], "20kb")
    compare_result_lists(inst.combined_scan(d, 500000), exp["combined_500kb"]["result"], "500kb")
    compare_result_lists(inst.scan_perChr_bySNPs(d, 500), exp["bysnps_500"]["result"], "500snps")


def test_packed_panel_fast_path_equals_dict_path(mods):
    """K0 (C++ packer) -> genotype-level GPU entry gives the same scans as the dict path on the same VCF."""
    K, _ = mods
    vcf, pm = os.path.join(GOLDEN, "ecb_subset.vcf.gz"), os.path.join(GOLDEN, "ecb_subset.popmap.txt")
    inst = K.LikelihoodInference_jointSFS(vcf, pm)
    d = inst.make_data_dict_vcf(vcf, pm)
    P = inst.make_packed_panel()
    assert len(P) == len(d)

    def same(a, b):
        assert list(a) == list(b)
        for k in a:
            for f in a[k]:
                assert close(a[k][f], b[k][f], 1e-10), (k, f, a[k][f], b[k][f])  # fused scan vs table scorer: different sum order

    rp = inst.combined_scan(P, 20000)
    assert inst._eng().h.scan_info() == (True, 4), "a PackedPanel scan must take the fused path with 4-byte records"
    same(rp, inst.combined_scan(d, 20000))
    assert inst._eng().h.scan_info()[0] is False  # the dict (counts) entry is scored by the table scorer
    rp = inst.scan_perChr_bySNPs(P, 500)
    assert inst._eng().h.scan_info()[0]
    same(rp, inst.scan_perChr_bySNPs(d, 500))
    same(inst.scan_chooseChr(P, 500000, "NC_087088.1"), inst.scan_chooseChr(d, 500000, "NC_087088.1"))
    assert inst.calculate_2d_sfs(P) == inst.calculate_2d_sfs(d)
    assert inst.calculate_1d_sfs(P, "bv", 14, None, None, None) == inst.calculate_1d_sfs(d, "bv", 14, None, None, None)
    bg2 = inst.calculate_2d_sfs(d)
    b1 = inst.fold_1d_sfs(inst.calculate_1d_sfs(d, "uv", 18, None, None, None))
    b2 = inst.fold_1d_sfs(inst.calculate_1d_sfs(d, "bv", 14, None, None, None))
    same(inst.scan_precomputed_BG(P, 500000, bg2, b1, b2), inst.scan_precomputed_BG(d, 500000, bg2, b1, b2))
    # the hand-built ingest fixture has half calls / haploid calls: fix-ups travel with the packed panel
    vcf, pm = os.path.join(GOLDEN, "ingest_small.vcf.gz"), os.path.join(GOLDEN, "ingest_small.popmap.txt")
    inst = K.LikelihoodInference_jointSFS(vcf, pm, pop1_size=3, pop2_size=3)
    d, P = inst.make_data_dict_vcf(vcf, pm), inst.make_packed_panel()
    assert P.fixups is not None and len(P.fixups) > 0
    assert inst.calculate_2d_sfs(P) == inst.calculate_2d_sfs(d)
    assert inst.calculate_1d_sfs(P, "uv", 3, None, None, None) == inst.calculate_1d_sfs(d, "uv", 3, None, None, None)


def _write_sim_vcf(path, rng, nsnp, L):
    import gzip
    samples = [f"i{k}" for k in range(10)]
    pos = np.sort(rng.choice(np.arange(1, L), size=nsnp, replace=False))
    with gzip.open(path, "wt") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n")
        for p in pos:
            fr = min(max(rng.beta(0.4, 0.8), 0.02), 0.98)
            gts = ["|".join(str(int(rng.random() < fr)) for _ in range(2)) for _ in samples]
            f.write(f"1\t{p}\t.\tA\tG\t.\tPASS\t.\tGT\t" + "\t".join(gts) + "\n")


def test_sims_directory_driver_batched(mods, tmp_path):
    """likelihood_scan over a simulation directory (BASELINE configs[2]): per-generation backgrounds from the concatenated
    VCF (1D left unfolded, region 0..500000), all replicates of a generation scored in one batched launch ==
    the oracle's process_window per replicate."""
    import sfs_oracle as O
    K, S = mods
    rng = np.random.default_rng(42)
    main = tmp_path / "sims"
    (main / "concatenated_vcfs").mkdir(parents=True)
    pm = tmp_path / "popmap.txt"
    pm.write_text("".join(f"i{k}\t{'p1' if k < 5 else 'p2'}\n" for k in range(10)))
    gens = ["1000", "2000"]
    for g in gens:
        _write_sim_vcf(str(main / "concatenated_vcfs" / f"gen.{g}.concatenated.vcf.gz"), rng, 3000, 1500000)
        for it in range(1, 4):
            d = main / f"iter{it}"
            d.mkdir(exist_ok=True)
            _write_sim_vcf(str(d / f"sim.{g}.{it}.vcf.gz"), rng, 1500, 1500000)
    assert S.get_gens(str(main)) == set(gens)
    got = S.likelihood_scan(str(main), popinfo_filename=str(pm))
    assert len(got) == 2 * 3 * 3
    for (g, it, key), rec in got.items():
        bgd = O.make_data_dict_vcf(str(main / "concatenated_vcfs" / f"gen.{g}.concatenated.vcf.gz"), str(pm))
        b2 = O.calculate_2d_sfs(bgd, "p1", "p2", 5, 5, 0, 500000, None)
        b1 = O.calculate_1d_sfs(bgd, "p1", 5, 0, 500000, None)
        b1b = O.calculate_1d_sfs(bgd, "p2", 5, 0, 500000, None)
        d = O.make_data_dict_vcf(str(main / f"iter{it}" / f"sim.{g}.{it}.vcf.gz"), str(pm))
        exp = O.sims_process_window(d, b2, b1, b1b, 500000, "p1", "p2", 5, 5, None, None, None)
        assert rec["region"] == ("background" if int(key.split("-")[1]) <= 1000000 else "foreground")
        for f, v in exp[key].items():
            assert close(rec["likelihood"][f], v), (g, it, key, f)
    out = tmp_path / "sims.csv"
    S.likelihood_scan_to_csv(str(main), str(out), popinfo_filename=str(pm))
    assert len(out.read_text().splitlines()) == 1 + 18
