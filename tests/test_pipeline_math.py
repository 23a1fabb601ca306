"""CPU check of the arithmetic behind the fused scan (csrc/tdsfs_fused.cuh: k1_fused + k3_finish): the statistic split into a
window-only part W = sum_bins x ln x (accumulated as dx[c] per insert, c = the bin's old count, in any arrival order) and a
background part G = sum_SNPs ln b[bin], with the per-bin form kept for one-bin windows and for N == B.  Compared with the
oracle's restatement of the reference likelihood (oracle/sfs_oracle.py clr_dense)."""
import math

import numpy as np

import sfs_oracle as O

LN = np.zeros(4096)
LN[1:] = np.log(np.arange(1, 4096, dtype=np.float64))
M = np.arange(4096, dtype=np.float64)
DX = np.zeros(4096)
DX[1:] = np.log(M[1:] + 1.0) + M[1:] * np.log1p(1.0 / M[1:])   # k_dx_table


def split_statistic(bins, b):
    """bins: interior bin index of every SNP of the window (arrival order); b: background counts per interior bin."""
    N = len(bins)
    B = float(b.sum())
    if N == 0 or B == 0:
        return None
    lb = np.where(b > 0, np.log(np.maximum(b, 1e-300)), -np.inf)
    seen = {}
    W = 0.0
    for k in bins:                       # k1_fused: dx of the old count
        c = seen.get(k, 0)
        W += DX[c]
        seen[k] = c + 1
    G = float(sum(lb[k] for k in bins))  # k3_finish
    acc = W - G
    if len(seen) == 1:                   # one populated bin: x (ln x - ln b) (exact_bins)
        k = bins[0]
        acc = float(N) * ((LN[N] if N > 1 else 0.0) - lb[k])
    if float(N) == B:                    # possibly its own background: per-bin form (exact_bins)
        acc = sum(float(x) * ((LN[x] if x > 1 else 0.0) - lb[k]) for k, x in seen.items())
    t = LN[N] - math.log(B)
    return 2.0 * (acc - float(N) * t)


def test_split_matches_reference_likelihood():
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(400):
        nb = int(rng.integers(2, 60))
        N = int(rng.integers(1, 700))
        p = rng.dirichlet(np.full(nb, 0.3))
        bins = rng.choice(nb, size=N, p=p)
        x = np.bincount(bins, minlength=nb)
        b = x + rng.integers(0, 50000, size=nb)          # the window is part of its background
        exp, none = O.clr_dense(x, b)
        got = split_statistic(list(rng.permutation(bins)), b.astype(np.float64))
        assert not none and got is not None
        worst = max(worst, abs(got - exp) / max(abs(exp), 1.0))
    assert worst <= 1e-9, worst


def test_split_exact_zeros_and_inf():
    # one bin in the window, the same single bin in the background, N != B
    assert split_statistic([3] * 17, np.array([0, 0, 0, 250.0, 0])) == 0.0
    assert O.clr_dense(np.array([0, 0, 0, 17, 0]), np.array([0, 0, 0, 250, 0]))[0] == 0.0
    # the window is the background
    rng = np.random.default_rng(5)
    bins = list(rng.integers(0, 30, size=500))
    x = np.bincount(bins, minlength=30).astype(np.float64)
    assert split_statistic(bins, x) == 0.0
    assert O.clr_dense(x.astype(np.int64), x)[0] == 0.0
    # a populated bin with an empty background bin: +inf like the reference
    b = x.copy()
    b[bins[0]] = 0
    assert split_statistic(bins, b) == math.inf
    assert O.clr_dense(x.astype(np.int64), b)[0] == math.inf
