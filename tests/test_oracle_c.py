"""Pin the C restatement (oracle/sfs_oracle.c) to the reference's golden chr1 outputs and to the Python oracle."""
import os
import subprocess

import numpy as np
import pytest

import sfs_oracle as O
from helpers import GOLDEN, load_chr1_arrays

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def OC():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    import sfs_oracle_c
    return sfs_oracle_c


def close_arr(a, b):
    return np.all(np.abs(a - b) <= 1e-9 * np.maximum(np.abs(b), 1))


@pytest.mark.parametrize("tag,kw", [("c20k", dict(W=20000)), ("c500k", dict(W=500000)), ("s500", dict(N=500))])
def test_c_oracle_chr1_golden(OC, tag, kw):
    chrom, pos, cnt, ann, vocab = load_chr1_arrays()
    runs = np.load(os.path.join(GOLDEN, "chr1_ref_runs.npz"))
    r = OC.scan(cnt, pos, [0, len(pos)], 18, 14, **kw)
    assert np.array_equal(r["start"], runs[f"{tag}_start"]) and np.array_equal(r["end"], runs[f"{tag}_end"])
    assert np.array_equal(r["snp_count"], runs[f"{tag}_snp_count"])
    for a, b in (("T2D", "T2D"), ("T1D_p1", "T1D_pop1"), ("T1D_p2", "T1D_pop2")):
        none = runs[f"{tag}_{b}_none"]
        assert np.array_equal(r[a + "_none"], none)
        assert close_arr(r[a][~none], runs[f"{tag}_{b}"][~none])


def test_c_oracle_matches_python_oracle_on_genotypes(OC):
    from tdsfs_pack import pack_codes
    rng = np.random.default_rng(11)
    S, n1, n2 = 5000, 40, 23
    c1 = rng.choice([0, 1, 3, 2], p=[0.7, 0.2, 0.07, 0.03], size=(S, n1)).astype(np.uint8)
    c2 = rng.choice([0, 1, 3, 2], p=[0.5, 0.3, 0.17, 0.03], size=(S, n2)).astype(np.uint8)
    G, w1, w2 = pack_codes(c1, c2)
    cnt = OC.decode(G, S, w1, w2, n1, n2)
    assert np.array_equal(cnt, O.unpack_counts(G, w1, w2, n1, n2, S))
    pos = np.sort(rng.choice(np.arange(0, 300000), size=S, replace=False)).astype(np.int32)
    off = np.array([0, 1200, 1200, S])
    for kw in (dict(W=9000), dict(N=130)):
        for bg in ("per_chrom", "genome"):
            a = OC.scan(cnt, pos, off, n1, n2, bg=bg, nthreads=2, **kw)
            b = O.scan_arrays(cnt, pos, off, n1, n2, bg=bg, **kw)
            for k in ("chrom", "start", "end", "snp_count"):
                assert np.array_equal(a[k], b[k]), k
            for k in ("T2D", "T1D_p1", "T1D_p2"):
                assert np.array_equal(a[k + "_none"], b[k + "_none"].astype(bool))
                ok = ~a[k + "_none"]
                assert close_arr(a[k][ok], b[k][ok])
