"""The fused scan (csrc/tdsfs_fused.cuh: k1_fused leaves every window's background-independent sums, k3_finish gathers
ln b over the per-SNP records) against the CPU oracle and against the table scorer on the same inputs: fixed-bp and
fixed-SNP windows, every background mode, per-SNP filter flags, chunked uploads (windows straddling launch boundaries),
windows above the warp-table capacity, the exact-0.0 / +inf / None cases, and the 4-byte -> 8-byte record fallback.
Integer work bit-exact, T2D / T1D within 1e-9 * max(|T|, 1) (north-star tolerance)."""
import os

import numpy as np
import pytest

import sfs_oracle as O
import sfs_oracle_c as OC
from test_gpu_capi_parity import compare_scan, random_panel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import tdsfs_capi
    return tdsfs_capi


@pytest.fixture()
def h(T):
    hd = T.Handle(0)
    yield hd
    hd.close()


def fused_scan(T, h, size, bg_mode, snp=False, bg_chrom=0, expect_fused=True, rec_bytes=None):
    h.plan(size, snp_mode=snp)
    h.background(bg_mode, bg_chrom)
    h.finalize_background()
    res = h.scan(size, snp_mode=snp)
    fused, rb = h.scan_info()
    assert fused == expect_fused, "the fused path did not run" if expect_fused else "the fused path ran unexpectedly"
    if rec_bytes is not None:
        assert rb == rec_bytes
    return res


def assert_same(a, b):
    for k in a:
        if a[k].dtype == np.float64:  # fp64 sums in a different order: equal to rounding, not bitwise
            fin = np.isfinite(b[k])
            assert np.array_equal(np.isfinite(a[k]), fin), k
            assert np.allclose(a[k][fin], b[k][fin], rtol=1e-10, atol=1e-10), k
            assert np.array_equal(a[k][~fin & ~np.isnan(b[k])], b[k][~fin & ~np.isnan(b[k])]), k
        else:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("n1,n2,S,C,L,W,N", [
    (18, 14, 60000, 3, 1200000, 20000, 250),     # ECB geometry (2 + 2 words)
    (5, 5, 90000, 40, 180000, 5000, 100),        # sims geometry, many chromosomes
    (200, 200, 120000, 5, 1200000, 20000, 500),  # config-4 geometry: k1_fused<14,14>
    (500, 500, 100000, 2, 2000000, 20000, 333),  # config-5 geometry: k1_fused<32,32>
    (100, 37, 30000, 2, 300000, 7000, 64),       # asymmetric, generic instantiation
    (33, 700, 20000, 3, 200000, 9000, 97),       # very asymmetric words (4 + 44)
])
@pytest.mark.parametrize("bg", ["per_chrom", "genome"])
def test_fused_vs_oracle(T, h, n1, n2, S, C, L, W, N, bg):
    rng = np.random.default_rng(n1 * 7919 + n2 * 13 + S)
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, C, L)
    cnt = OC.decode(G, S, w1, w2, n1, n2, nthreads=4)
    mode = T.BG_PER_CHROM if bg == "per_chrom" else T.BG_GENOME
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    res = fused_scan(T, h, W, mode, rec_bytes=4)
    compare_scan(T, res, OC.scan(cnt, pos, off, n1, n2, W=W, bg=bg, nthreads=4))
    # the table scorer on the same records (scan without a plan) agrees
    h.background(mode)
    h.finalize_background()
    plain = h.scan(W)
    assert h.scan_info()[0] is False
    assert_same(res, plain)
    # fixed-SNP windows
    res = fused_scan(T, h, N, mode, snp=True)
    compare_scan(T, res, OC.scan(cnt, pos, off, n1, n2, N=N, bg=bg, nthreads=4), snp_mode=True)


def test_fused_small_python_oracle_all_quirks(T, h):
    """Small case against the pure-Python oracle: sparse windows (one-bin windows, empty windows, None statistics)."""
    rng = np.random.default_rng(77)
    n1, n2, S = 6, 4, 3000
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 6, 400000, miss=0.1)
    cnt = O.unpack_counts(G, w1, w2, n1, n2, S)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    for W in (300, 2000, 50000):
        for mode, bg in ((T.BG_PER_CHROM, "per_chrom"), (T.BG_GENOME, "genome")):
            compare_scan(T, fused_scan(T, h, W, mode), O.scan_arrays(cnt, pos, off, n1, n2, W=W, bg=bg))
    for N in (1, 7, 50):
        compare_scan(T, fused_scan(T, h, N, T.BG_PER_CHROM, snp=True), O.scan_arrays(cnt, pos, off, n1, n2, N=N), snp_mode=True)


def test_fused_exact_zero_self_background_and_inf(T, h):
    """A window that is its own background scores exactly 0.0 (the reference's truthiness quirk depends on it); a window of a
    chromosome outside a single-chromosome background has bins with zero background: +inf."""
    rng = np.random.default_rng(5)
    n1, n2, S = 20, 12, 1500
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 3, 5000)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    res = fused_scan(T, h, 1000000, T.BG_PER_CHROM)  # one window per chromosome == its background
    live = (res["flags"] & T.F_EMPTY) == 0
    assert live.sum() == 3
    for k in ("T2D", "T1D_p1", "T1D_p2"):
        assert np.all(res[k][live] == 0.0), (k, res[k][live])
    res = fused_scan(T, h, 500, T.BG_CHROM, bg_chrom=1)
    plain_h = T.Handle(0)
    plain_h.set_panel(n1, n2, True)
    plain_h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    plain_h.background(T.BG_CHROM, 1)
    plain_h.finalize_background()
    plain = plain_h.scan(500)
    plain_h.close()
    assert np.isinf(res["T2D"]).any()
    assert_same(res, plain)


def test_fused_flags_and_large_windows(T, h):
    """Per-SNP filter flags (bit0 spectrum filter, bit1 count_snps) and windows above WCAP (CTA path inside k3_finish)."""
    rng = np.random.default_rng(31)
    n1, n2, S = 30, 40, 60000
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 2, 1200000)
    flags = (rng.random(S) < 0.8).astype(np.uint8) | ((rng.random(S) < 0.7).astype(np.uint8) << 1)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off, flags=flags)
    h2 = T.Handle(0)
    h2.set_panel(n1, n2, True)
    h2.load_genotypes(G, S, w1, w2, n1, n2, pos, off, flags=flags)
    for W, snp in ((15000, False), (100000, False), (400, True), (3000, True)):
        res = fused_scan(T, h, W, T.BG_GENOME, snp=snp, expect_fused=not (snp and W > 768))
        h2.background(T.BG_GENOME)
        h2.finalize_background()
        plain = h2.scan(W, snp_mode=snp)
        assert W != 100000 or res["snp_count"].max() > 768
        assert_same(res, plain)
    h2.close()
    cnt = O.unpack_counts(G, w1, w2, n1, n2, S)
    res = fused_scan(T, h, 15000, T.BG_PER_CHROM)
    exp = OC.scan(cnt, pos, off, n1, n2, W=15000, bg="per_chrom", nthreads=4, include_flags=(flags & 1))
    live = (res["flags"] & T.F_EMPTY) == 0
    assert np.array_equal(res["start"][live], exp["start"])
    for a, bit in (("T2D", T.F_T2D_NONE), ("T1D_p1", T.F_T1D_P1_NONE), ("T1D_p2", T.F_T1D_P2_NONE)):
        none = (res["flags"][live] & bit) != 0
        assert np.array_equal(none, exp[a + "_none"])
        err = np.abs(res[a][live][~none] - exp[a][~none]) / np.maximum(np.abs(exp[a][~none]), 1.0)
        assert err.max() <= 1e-9, (a, err.max())


def test_fused_chunked_upload_windows_straddle_launches(T, monkeypatch):
    """Host matrix uploaded in many small chunks: every chunk is one k1_fused launch and windows straddle the boundaries."""
    monkeypatch.setenv("TDSFS_UPLOAD_CHUNK_KB", "64")
    rng = np.random.default_rng(8)
    n1, n2, S = 200, 200, 50000
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 3, 600000)
    cnt = OC.decode(G, S, w1, w2, n1, n2, nthreads=4)
    h = T.Handle(0)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    res = h.run_bp(T.BG_GENOME, 20000)
    assert h.scan_info()[0]
    compare_scan(T, res, OC.scan(cnt, pos, off, n1, n2, W=20000, bg="genome", nthreads=4))
    res = fused_scan(T, h, 123, T.BG_PER_CHROM, snp=True)
    compare_scan(T, res, OC.scan(cnt, pos, off, n1, n2, N=123, bg="per_chrom", nthreads=4), snp_mode=True)
    h.close()


def test_narrow_record_overflow_falls_back_to_wide(T, h):
    """More missing calls than the 4-byte record's field holds: run_bp retries with 8-byte records by itself."""
    from tdsfs_pack import pack_codes
    rng = np.random.default_rng(3)
    n1 = n2 = 500
    S = 4000
    f = rng.uniform(0.55, 0.95, size=S)[:, None]  # high alt frequency: swapped SNPs

    def codes(ns):
        a = (rng.random((S, ns)) < f).astype(np.uint8) + (rng.random((S, ns)) < f).astype(np.uint8)
        c = np.where(a == 2, 3, a).astype(np.uint8)
        c[rng.random((S, ns)) < 0.3] = 2  # 30 % missing: ~150 missing diploids per population > 63
        return c

    G, w1, w2 = pack_codes(codes(n1), codes(n2))
    pos = np.sort(rng.choice(np.arange(1, 400000), size=S, replace=False)).astype(np.int32)
    off = np.array([0, S], dtype=np.int64)
    cnt = OC.decode(G, S, w1, w2, n1, n2, nthreads=4)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    res = h.run_bp(T.BG_GENOME, 20000)
    assert h.scan_info() == (True, 8)
    compare_scan(T, res, OC.scan(cnt, pos, off, n1, n2, W=20000, bg="genome", nthreads=4))
    # synchronous step-by-step calls fall back inside tdsfs_background
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    res2 = fused_scan(T, h, 20000, T.BG_GENOME, rec_bytes=8)
    assert_same(res2, res)


def test_fused_many_warp_ranges_and_positions_without_tma(T, monkeypatch):
    """Enough SNPs that every warp of every CTA owns a range (1184+ ranges), once with the positions through the TMA ring and
    once through plain loads (the fallback for position arrays that are not 16-byte aligned)."""
    rng = np.random.default_rng(12)
    n1, n2, S = 64, 64, 400000
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 7, 4000000, miss=0.01)
    cnt = OC.decode(G, S, w1, w2, n1, n2, nthreads=8)
    exp = OC.scan(cnt, pos, off, n1, n2, W=10000, bg="genome", nthreads=8)
    for no_tma in ("", "1"):
        if no_tma:
            monkeypatch.setenv("TDSFS_NO_POS_TMA", "1")
        h = T.Handle(0)
        h.set_panel(n1, n2, True)
        h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
        res = fused_scan(T, h, 10000, T.BG_GENOME)
        compare_scan(T, res, exp)
        live = (res["flags"] & T.F_EMPTY) == 0
        assert int(res["snp_count"][live].sum()) == S
        h.close()


def test_fused_float_background(T, h):
    """scan_precomputed_BG / sims usage: keys only, caller-supplied float background, planned scan."""
    rng = np.random.default_rng(21)
    n1, n2, S = 18, 14, 20000
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 2, 400000)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off)
    h.background(T.BG_GENOME)
    s2, s1a, s1b = h.get_background(0)
    b2 = s2.astype(np.float64)
    b2 /= b2.ravel()[1:-1].sum()
    f1, f2 = O.fold_dense(s1a.astype(np.int64)).astype(np.float64), O.fold_dense(s1b.astype(np.int64)).astype(np.float64)
    f1 /= f1[1:-1].sum()
    f2 /= f2[1:-1].sum()
    out = {}
    for planned in (False, True):
        if planned:
            h.plan(20000)
        h.background(T.BG_NONE)
        h.set_background(b2, f1, f2)
        out[planned] = h.scan(20000)
        assert h.scan_info()[0] == planned
    assert_same(out[True], out[False])


def test_step_bp_graph_replay_equals_run_bp(T, h):
    """tdsfs_step_bp on device-resident data: eager, captured and replayed passes leave the same results as tdsfs_run_bp;
    other arguments or a new load drop the graph."""
    import torch
    rng = np.random.default_rng(41)
    n1, n2, S = 64, 64, 60000
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 3, 900000)
    Gd, pd = torch.from_numpy(G.view(np.int32)).cuda(), torch.from_numpy(pos).cuda()
    h.set_panel(n1, n2, True)
    h.load_genotypes(Gd, S, w1, w2, n1, n2, pd, off)
    ref = h.run_bp(T.BG_GENOME, 15000)
    h.set_sync(False)
    l0 = h.launch_count()
    per_pass = []
    for i in range(5):  # 1: eager, 2: capture + launch, 3..: replay
        h.step_bp(T.BG_GENOME, 15000)
        h.check()
        per_pass.append(h.launch_count() - l0)
        l0 = h.launch_count()
        assert_same(h.fetch_results(len(ref["start"])), ref)
    assert len(set(per_pass)) == 1, per_pass  # the replayed graph accounts for the same kernels as the eager pass
    h.step_bp(T.BG_PER_CHROM, 15000)  # other arguments: eager again
    h.check()
    got = h.fetch_results(len(ref["start"]))
    h.set_sync(True)
    assert_same(got, h.run_bp(T.BG_PER_CHROM, 15000))
    G2, w1, w2, pos2, off2 = random_panel(rng, S // 2, n1, n2, 2, 500000)
    Gd2, pd2 = torch.from_numpy(G2.view(np.int32)).cuda(), torch.from_numpy(pos2).cuda()
    h.load_genotypes(Gd2, S // 2, w1, w2, n1, n2, pd2, off2)  # new data: the old graph must not be replayed
    h.set_sync(False)
    for i in range(3):
        h.step_bp(T.BG_GENOME, 15000)
    h.check()
    got = h.fetch_results(h.candidates(15000))
    h.set_sync(True)
    assert_same(got, h.run_bp(T.BG_GENOME, 15000))
