"""Legacy Poisson window scan on the GPU (tdsfs_set_poisson_background + tdsfs_scan_poisson_bp, drop-in module twoDSFS.py)
against outputs of the UNMODIFIED first-generation script (tests/golden/poisson_cases.json) and against the oracle on
larger seeded inputs (windows above the warp-table capacity included).  Tolerance 1e-9 * max(|P|, 1)."""
import json
import os

import numpy as np
import pytest

import sfs_oracle as O
from helpers import GOLDEN, close, rows_to_dict

pytestmark = pytest.mark.gpu


def test_drop_in_matches_reference_runs():
    import twoDSFS as L
    for c in json.load(open(os.path.join(GOLDEN, "poisson_cases.json")))["cases"]:
        d = rows_to_dict(c["rows"], ("uv", "bv"))
        src = d if c["bg_rows"] is None else rows_to_dict(c["bg_rows"], ("uv", "bv"))
        bg = L.calculate_2d_sfs(src, "uv", "bv", c["n1"], c["n2"], None, None, None)
        bgn = L.normalize_2d_sfs(bg)
        for (i, j, v) in c["bg_norm"]:
            assert close(bgn[(i, j)], v, 1e-12)
        got = L.calculate_p_window(d, bgn, c["W"], "uv", "bv", c["n1"], c["n2"], c["start"], c["end"], c["variant_type"])
        assert list(got.keys()) == [w[0] for w in c["windows"]]
        for key, p, cnt in c["windows"]:
            assert got[key]["snp_count"] == cnt
            assert close(got[key]["p_values"], p), (key, got[key]["p_values"], p)
        # calculate_p on explicit spectra (one CTA kernel) agrees with the window scan's first window
        first = c["windows"][0][0]
        chrom, rng_ = first.split(" ")
        a, b = (int(v) for v in rng_.split("-"))
        win = {k: v for k, v in d.items() if k.split("-")[0] == chrom and a <= max(int(k.split("-")[1]), 1) <= b}
        fg = L.calculate_2d_sfs(win, "uv", "bv", c["n1"], c["n2"], c["start"], c["end"], c["variant_type"])
        assert close(L.calculate_p(fg, bgn), c["windows"][0][1])


def test_capi_poisson_scan_vs_oracle_large_windows():
    """Genotype entry, 40 + 30 diploids, windows of ~300 and ~3000 SNPs (CTA path), filter flags."""
    import tdsfs_capi as T
    from test_gpu_capi_parity import random_panel
    rng = np.random.default_rng(61)
    n1, n2, S = 40, 30, 30000
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 2, 600000)
    cnt = O.unpack_counts(G, w1, w2, n1, n2, S)
    flags = (rng.random(S) < 0.9).astype(np.uint8) | ((rng.random(S) < 0.8).astype(np.uint8) << 1)
    R1, R2 = 2 * n1 + 1, 2 * n2 + 1
    inc = (flags & 1) != 0
    bg = np.bincount(cnt[inc, 1] * R2 + cnt[inc, 3], minlength=R1 * R2).astype(np.float64)
    bg[0] = 0
    tot = bg.sum()
    bg += 1.0 / tot
    q = bg / bg[1:-1].sum()
    h = T.Handle(0)
    h.set_panel(n1, n2, False)
    h.load_genotypes(G, S, w1, w2, n1, n2, pos, off, flags=flags)
    h.background(T.BG_NONE)
    h.set_poisson_background(q)
    for W in (12000, 120000):
        res = h.scan_poisson(W)
        live = np.flatnonzero((res["flags"] & T.F_EMPTY) == 0)
        assert W != 120000 or res["snp_count"].max() > 768
        for wid in live[:: max(1, len(live) // 12)].tolist():
            c = int(res["chrom"][wid])
            pc = pos[off[c]:off[c + 1]]
            lo = int(off[c] + np.searchsorted(np.maximum(pc, 1), res["start"][wid], side="left"))
            hi = int(off[c] + np.searchsorted(np.maximum(pc, 1), res["end"][wid], side="right"))
            sel = np.arange(lo, hi)[inc[lo:hi]]
            x = np.bincount(cnt[sel, 1] * R2 + cnt[sel, 3], minlength=R1 * R2)
            x[0] = 0
            assert res["n2d"][wid] == int(x.sum())
            assert res["snp_count"][wid] == int(((flags[lo:hi] >> 1) & 1).sum())
            exp = O.poisson_window_score(x, q)
            assert close(res["T2D"][wid], exp), (W, wid, res["T2D"][wid], exp)
    with pytest.raises(T.TdsfsError):
        h.scan(12000)          # a likelihood scan against Poisson tables is refused
    h.close()
