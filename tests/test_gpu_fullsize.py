"""BASELINE.json full-size synthetic configuration (configs[3]: 10 M SNPs x 200+200 diploids, 20 kb windows) through
size-independent properties, plus sampled windows against the C oracle:
  * every SNP is in exactly one window: sum(snp_count) == S, windows ordered, boundaries aligned to 1 + k*W
  * background spectrum == sum of the per-chromosome background spectra (linearity), total == number of unskipped SNPs
  * the genome-wide scan's window spectra sum to the background (checksum of checksums on sampled chromosomes)
  * T2D / T1D of 1,000 seeded windows == oracle on the same rows scored against the GPU background (1e-9)
The 50 M x 1000 configuration runs in bench.py; this keeps the test suite at a few seconds of GPU time."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config4_full_size_properties():
    import torch
    import tdsfs_capi as T
    import sfs_oracle as O
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.WORKLOADS["config4"]
    S, n1, n2, W, C = cfg["S"], cfg["n1"], cfg["n2"], cfg["W"], cfg["C"]
    w1, w2 = bench.words_for(n1), bench.words_for(n2)
    RW = w1 + w2
    pos = np.concatenate(bench.positions_for(cfg, range(C)))
    off = np.concatenate([[0], np.cumsum(bench.chrom_sizes(S, C))]).astype(np.int64)
    h = T.Handle(0)
    h.set_panel(n1, n2, True)
    G = torch.empty(((S + 31) // 32 * RW * 32,), dtype=torch.int32, device="cuda")
    h.synth_genotypes(G.data_ptr(), S, 0, w1, w2, n1, n2, cfg["seed"])
    pos_dev = torch.from_numpy(pos).cuda()
    h.load_genotypes(G, S, w1, w2, n1, n2, pos_dev, off)

    # genome-wide background and scan
    res = h.run_bp(T.BG_GENOME, W)
    g2, g1a, g1b = h.get_background(0)
    live = (res["flags"] & T.F_EMPTY) == 0
    assert int(res["snp_count"][live].sum()) == S
    assert np.all((res["start"] - 1) % W == 0) and np.all(res["end"] - res["start"] == W - 1)
    key = res["chrom"].astype(np.int64) * (1 << 40) + res["start"]
    assert np.all(np.diff(key) > 0)
    assert int(g2.sum()) == int(res["n2d"][live].sum()) + int(g2[-1, -1])  # every unskipped SNP is in exactly one window
    # folded 1D interiors: background total == sum over windows
    f1 = O.fold_dense(g1a.astype(np.int64))
    assert int(f1[1:-1].sum()) == int(res["n1d_p1"][live].sum())
    T2_genome = res["T2D"].copy()

    # linearity: per-chromosome backgrounds sum to the genome-wide one
    h.background(T.BG_PER_CHROM)
    acc2 = np.zeros_like(g2)
    acc1 = np.zeros_like(g1a)
    for c in range(C):
        s2, s1a, _ = h.get_background(c)
        acc2 += s2
        acc1 += s1a
    assert np.array_equal(acc2, g2) and np.array_equal(acc1, g1a)

    # 1,000 sampled windows vs the oracle (SURVEY.md 8(d) parity at scale), scored against the GPU's genome-wide background;
    # the window's rows are copied back from HBM and decoded by the C oracle
    import sfs_oracle_c as OC
    h.plan(W)
    h.background(T.BG_GENOME)
    h.finalize_background()
    h.scan(W, fetch=False)
    assert h.scan_info()[0], "the fused path must score the full-size configuration"
    rng = np.random.default_rng(3)
    ids = rng.choice(np.flatnonzero(live), size=1000, replace=False)
    b2 = g2.astype(np.int64).ravel()[1:-1]
    b1a, b1b = O.fold_dense(g1a.astype(np.int64))[1:-1], O.fold_dense(g1b.astype(np.int64))[1:-1]
    cand = np.concatenate([[0], np.cumsum([(pos[off[c + 1] - 1] - 1) // W + 1 for c in range(C)])])
    worst = 0.0
    for n_done, wid in enumerate(ids.tolist()):
        c = int(res["chrom"][wid])
        lo = off[c] + np.searchsorted(pos[off[c]:off[c + 1]], res["start"][wid], side="left")
        hi = off[c] + np.searchsorted(pos[off[c]:off[c + 1]], res["end"][wid], side="right")
        assert hi - lo == res["snp_count"][wid] and cand[c] <= wid < cand[c + 1]
        blk0, blk1 = lo // 32, (hi + 31) // 32
        words = G[blk0 * RW * 32:blk1 * RW * 32].cpu().numpy().view(np.uint32)
        cnt = OC.decode(words, (blk1 - blk0) * 32, w1, w2, n1, n2)[lo - blk0 * 32:hi - blk0 * 32]
        e2, e1, e1b = O.dense_spectra(cnt, n1, n2)
        if n_done < 24:  # dense window spectra through the C ABI (one launch + a 640 KB copy each): a subset
            s2, s1a, s1b = h.window_spectra(wid)
            assert np.array_equal(s2.astype(np.int64), e2) and np.array_equal(s1a.astype(np.int64), e1) and np.array_equal(s1b.astype(np.int64), e1b)
        for name, x, b in (("T2D", e2.ravel()[1:-1], b2), ("T1D_p1", O.fold_dense(e1)[1:-1], b1a), ("T1D_p2", O.fold_dense(e1b)[1:-1], b1b)):
            exp, none = O.clr_dense(x, b)
            assert not none
            got = res[name][wid]
            err = abs(got - exp) / max(abs(exp), 1.0)
            worst = max(worst, err)
            assert err <= 1e-9, (wid, name, got, exp)
    print("config 4: 1000 windows vs the oracle, max relative error %.3e" % worst)
    assert np.allclose(h.fetch_results(len(res["start"]))["T2D"][live], T2_genome[live], rtol=1e-11, atol=1e-11)
    h.close()
