"""Every instantiation of the shared-memory window scorer (k3_score_small with 1, 2 and 4 warps per window) against the
CPU oracle on the same inputs.  The library picks the group width from the scan size; the tuning knob TDSFS_SCORE_G
(read at every scan) forces one so that all of them are parity-tested on small inputs.  Includes the spectra for which
the reference returns exactly 0.0 (its truthiness drives the stale-carry quirk): a window that is its
own background, and a window whose only bin is the background's only bin."""
import os

import numpy as np
import pytest

import sfs_oracle as O
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


@pytest.fixture()
def knobs():
    saved = {k: os.environ.get(k) for k in ("TDSFS_SCORE_G",)}

    def set_(G):
        os.environ["TDSFS_SCORE_G"] = str(G)
    yield set_
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("G", [1, 2, 4])
@pytest.mark.parametrize("n1,n2,S,C,L,W,N", [
    (18, 14, 12000, 3, 300000, 20000, 250),    # ECB geometry, ~800 SNPs per window: small and large windows mixed
    (200, 37, 9000, 2, 100000, 6000, 300),     # asymmetric panel, ~270 SNPs per window
])
def test_variant_vs_oracle(T, h, knobs, G, n1, n2, S, C, L, W, N):
    rng = np.random.default_rng(n1 * 7 + n2 + S)
    Gm, w1, w2, pos, off = random_panel(rng, S, n1, n2, C, L)
    cnt = O.unpack_counts(Gm, w1, w2, n1, n2, S)
    knobs(G)
    h.set_panel(n1, n2, True)
    h.load_genotypes(Gm, S, w1, w2, n1, n2, pos, off)
    for bg in ("per_chrom", "genome"):
        h.background(T.BG_PER_CHROM if bg == "per_chrom" else T.BG_GENOME)
        h.finalize_background()
        compare_scan(T, h.scan(W), O.scan_arrays(cnt, pos, off, n1, n2, W=W, bg=bg))
        compare_scan(T, h.scan(N, snp_mode=True), O.scan_arrays(cnt, pos, off, n1, n2, N=N, bg=bg), snp_mode=True)


@pytest.mark.parametrize("G", [1, 2])
def test_variant_flags_and_inf(T, h, knobs, G):
    """Per-SNP flags (spectrum filter bit 0, count_snps filter bit 1) and a precomputed background with empty bins (+inf)."""
    rng = np.random.default_rng(5)
    n1, n2, S = 12, 9, 6000
    Gm, w1, w2, pos, off = random_panel(rng, S, n1, n2, 2, 120000)
    cnt = O.unpack_counts(Gm, w1, w2, n1, n2, S)
    flags = rng.integers(0, 4, size=S).astype(np.uint8)
    knobs(G)
    h.set_panel(n1, n2, True)
    h.load_counts(cnt.astype(np.uint16), pos, off, flags=flags)
    h.background(T.BG_PER_CHROM)
    h.finalize_background()
    res = h.scan(5000)
    exp = O.scan_arrays(cnt, pos, off, n1, n2, W=5000, include=(flags & 1).astype(bool))
    live = (res["flags"] & T.F_EMPTY) == 0
    for k in ("chrom", "start", "end"):
        assert np.array_equal(res[k][live], exp[k]), k
    for a, bit in (("T2D", T.F_T2D_NONE), ("T1D_p1", T.F_T1D_P1_NONE), ("T1D_p2", T.F_T1D_P2_NONE)):
        none = (res["flags"][live] & bit) != 0
        assert np.array_equal(none, exp[a + "_none"].astype(bool)), a
        d = np.abs(res[a][live][~none] - exp[a][~none]) / np.maximum(np.abs(exp[a][~none]), 1.0)
        assert d.size == 0 or d.max() <= 1e-9, (a, d.max())
    # background = spectra of chromosome 0 only: windows of chromosome 1 hit empty background bins -> +inf like the reference
    h.load_counts(cnt.astype(np.uint16), pos, off)
    h.background(T.BG_CHROM, bg_chrom=0)
    h.finalize_background()
    res = h.scan(5000)
    live = (res["flags"] & T.F_EMPTY) == 0
    b2, b1, b1b = O.dense_spectra(cnt[off[0]:off[1]], n1, n2)
    wins = list(O.bp_window_ranges(pos, off, 5000))
    assert int(live.sum()) == len(wins)
    got = {k: res[k][live] for k in ("T2D", "T1D_p1", "T1D_p2")}
    n_inf = 0
    for i, (c, s, lo, hi) in enumerate(wins):
        h2, h1, h1b = O.dense_spectra(cnt[lo:hi], n1, n2)
        for name, x, b in (("T2D", h2.ravel()[1:-1], b2.ravel()[1:-1]), ("T1D_p1", O.fold_dense(h1)[1:-1], O.fold_dense(b1)[1:-1]),
                           ("T1D_p2", O.fold_dense(h1b)[1:-1], O.fold_dense(b1b)[1:-1])):
            e, none = O.clr_dense(x, b)
            if none:
                continue
            g = got[name][i]
            if np.isinf(e):
                n_inf += 1
                assert g == e, (name, i, g, e)
            else:
                assert abs(g - e) <= 1e-9 * max(abs(e), 1.0), (name, i, g, e)
    assert n_inf > 0


@pytest.mark.parametrize("G", [1, 2, 4])
def test_variant_exact_zeros(T, h, knobs, G):
    """Exactly 0.0, bit for bit, where the reference's p_fg == p_bg: (a) windows that are their own background (one window
    per chromosome, per-chromosome background), (b) every SNP in one bin (one-bin background, many windows with N != B)."""
    rng = np.random.default_rng(17)
    n1, n2 = 20, 11
    knobs(G)
    h.set_panel(n1, n2, True)
    # (a) 5 chromosomes of 60..700 SNPs, one 1 Mb window each
    sizes = [60, 700, 333, 1, 512]
    S = sum(sizes)
    Gm, w1, w2, _, _ = random_panel(rng, S, n1, n2, 1, 900000)
    cnt = O.unpack_counts(Gm, w1, w2, n1, n2, S)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pos = np.concatenate([np.sort(rng.choice(np.arange(1, 900000), size=s, replace=False)) for s in sizes]).astype(np.int32)
    h.load_counts(cnt.astype(np.uint16), pos, off)
    h.background(T.BG_PER_CHROM)
    h.finalize_background()
    res = h.scan(1000000)
    exp = O.scan_arrays(cnt, pos, off, n1, n2, W=1000000)
    compare_scan(T, res, exp)
    live = (res["flags"] & T.F_EMPTY) == 0
    for a, bit in (("T2D", T.F_T2D_NONE), ("T1D_p1", T.F_T1D_P1_NONE), ("T1D_p2", T.F_T1D_P2_NONE)):
        none = (res["flags"][live] & bit) != 0
        assert np.all(exp[a][~none] == 0.0)
        assert np.all(res[a][live][~none] == 0.0), (a, res[a][live][~none])   # exactly zero, not 1e-13
    # (b) 3000 identical SNPs: ref/alt = (30, 10) and (19, 3) -> one 2D bin, one 1D bin per population
    S = 3000
    cnt = np.tile(np.array([[30, 10, 19, 3]], dtype=np.uint16), (S, 1))
    pos = np.sort(rng.choice(np.arange(1, 300000), size=S, replace=False)).astype(np.int32)
    off = np.array([0, S], dtype=np.int64)
    h.load_counts(cnt, pos, off)
    h.background(T.BG_GENOME)
    h.finalize_background()
    res = h.scan(20000)
    exp = O.scan_arrays(cnt, pos, off, n1, n2, W=20000, bg="genome")
    compare_scan(T, res, exp)
    live = (res["flags"] & T.F_EMPTY) == 0
    assert live.sum() > 5 and not (res["flags"][live] & 7).any()
    for a in ("T2D", "T1D_p1", "T1D_p2"):
        assert np.all(exp[a] == 0.0) and np.all(res[a][live] == 0.0), (a, res[a][live])


def test_group_widths_agree_to_rounding(T, h, knobs):
    """1, 2 and 4 warps per window compute the same sums in different orders: equal to rounding on the same scan."""
    rng = np.random.default_rng(23)
    n1, n2, S = 64, 64, 30000
    Gm, w1, w2, pos, off = random_panel(rng, S, n1, n2, 4, 300000)
    h.set_panel(n1, n2, True)
    h.load_genotypes(Gm, S, w1, w2, n1, n2, pos, off)
    h.background(T.BG_GENOME)
    h.finalize_background()
    out = {}
    for G in (1, 2, 4):
        knobs(G)
        out[G] = h.scan(10000)
    for G in (2, 4):
        for k, v in out[1].items():
            if v.dtype == np.float64:
                assert np.allclose(out[G][k], v, rtol=1e-11, atol=1e-11, equal_nan=True), (G, k)
            else:
                assert np.array_equal(out[G][k], v), (G, k)
