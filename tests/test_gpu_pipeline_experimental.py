"""EXPERIMENTAL, opt-in (TDSFS_TEST_PIPELINE=1): the pipelined scorer of csrc/tdsfs_pipeline.cuh (TDSFS_PIPELINE=1: window
sums under the count kernel on a second stream, gather-and-finish after the background) against the default table-walk
scorer and the CPU oracle.  Skipped by default: the code path is off in the product until it has been measured and this
test has passed on a B200."""
import os

import numpy as np
import pytest

import sfs_oracle as O
from test_gpu_capi_parity import compare_scan, random_panel

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TDSFS_TEST_PIPELINE") != "1", reason="experimental path: opt in with TDSFS_TEST_PIPELINE=1")]


def _handle(pipeline, chunks=3):
    import tdsfs_capi as T
    saved = {k: os.environ.get(k) for k in ("TDSFS_PIPELINE", "TDSFS_PIPELINE_CHUNKS")}
    os.environ["TDSFS_PIPELINE"] = "1" if pipeline else "0"   # read by tdsfs_create
    os.environ["TDSFS_PIPELINE_CHUNKS"] = str(chunks)         # read by tdsfs_background
    try:
        return T, T.Handle(0)
    finally:
        if saved["TDSFS_PIPELINE"] is None:
            os.environ.pop("TDSFS_PIPELINE", None)
        else:
            os.environ["TDSFS_PIPELINE"] = saved["TDSFS_PIPELINE"]


@pytest.mark.parametrize("n1,n2,S,C,L,W", [(18, 14, 20000, 3, 400000, 10000), (200, 200, 40000, 5, 400000, 20000),
                                           (500, 500, 12000, 2, 200000, 20000)])
def test_pipeline_vs_oracle_and_default(n1, n2, S, C, L, W):
    import torch
    rng = np.random.default_rng(n1 + n2 + S)
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, C, L)
    cnt = O.unpack_counts(G, w1, w2, n1, n2, S)
    exp = O.scan_arrays(cnt, pos, off, n1, n2, W=W, bg="genome")
    Gd = torch.from_numpy(G.view(np.int32)).cuda()       # device-resident matrix: the pipelined path needs one chunk
    pd = torch.from_numpy(pos).cuda()
    out = {}
    for pipeline in (False, True):
        T, h = _handle(pipeline)
        h.set_panel(n1, n2, True)
        h.load_genotypes(Gd, S, w1, w2, n1, n2, pd, off)
        launches0 = h.launch_count()
        out[pipeline] = h.run_bp(T.BG_GENOME, W)
        nl = h.launch_count() - launches0
        compare_scan(T, out[pipeline], exp)
        if pipeline:
            assert nl >= 3 * 2 + 3, nl   # K2, 3 x (K1 + window sums), finalize, gather, large: the path really ran
        h.close()
    for k, v in out[False].items():
        if v.dtype == np.float64:
            assert np.allclose(out[True][k], v, rtol=1e-11, atol=1e-11, equal_nan=True), k
        else:
            assert np.array_equal(out[True][k], v), k


def test_pipeline_exact_zeros():
    """One-bin windows (exact 0.0 where the background has the same single bin) and a window that is the whole background."""
    import torch
    from tdsfs_pack import pack_codes
    T, h = _handle(True, chunks=2)
    n1, n2, S = 20, 11, 4000
    rng = np.random.default_rng(3)
    # every SNP identical: 10 het samples in pop 1, 3 in pop 2 -> one 2D bin, one 1D bin per population
    c1 = np.zeros((S, n1), dtype=np.uint8); c1[:, :10] = 1
    c2 = np.zeros((S, n2), dtype=np.uint8); c2[:, :3] = 1
    G, w1, w2 = pack_codes(c1, c2)
    pos = np.sort(rng.choice(np.arange(1, 400000), size=S, replace=False)).astype(np.int32)
    off = np.array([0, S], dtype=np.int64)
    cnt = O.unpack_counts(G, w1, w2, n1, n2, S)
    h.set_panel(n1, n2, True)
    h.load_genotypes(torch.from_numpy(G.view(np.int32)).cuda(), S, w1, w2, n1, n2, torch.from_numpy(pos).cuda(), off)
    for W in (20000, 1000000):   # many one-bin windows; then one window == the background (N == B; > 768 SNPs -> CTA scorer)
        res = h.run_bp(T.BG_GENOME, W)
        exp = O.scan_arrays(cnt, pos, off, n1, n2, W=W, bg="genome")
        compare_scan(T, res, exp)
        live = (res["flags"] & T.F_EMPTY) == 0
        for a in ("T2D", "T1D_p1", "T1D_p2"):
            assert np.all(exp[a] == 0.0) and np.all(res[a][live] == 0.0), (W, a, res[a][live])
    # a small window (600 SNPs, many bins) that is the whole background: N == B -> re-scored per bin, exactly 0.0
    S = 600
    G, w1, w2, pos, off = random_panel(rng, S, n1, n2, 1, 500000)
    cnt = O.unpack_counts(G, w1, w2, n1, n2, S)
    h.load_genotypes(torch.from_numpy(G.view(np.int32)).cuda(), S, w1, w2, n1, n2, torch.from_numpy(pos).cuda(), off)
    res = h.run_bp(T.BG_GENOME, 1000000)
    compare_scan(T, res, O.scan_arrays(cnt, pos, off, n1, n2, W=1000000, bg="genome"))
    live = (res["flags"] & T.F_EMPTY) == 0
    assert live.sum() == 1 and all(res[a][live][0] == 0.0 for a in ("T2D", "T1D_p1", "T1D_p2"))
    h.close()
