"""Randomised differential test: the drop-in Python API (GPU) against the pinned CPU oracle on seeded random data_dicts
that hit the quirk ledger (None windows, stale-carry, first/last-window exceptions, +inf, filters, missing populations,
position 0, string-ordered chromosomes).  Results AND exception types must agree."""
import numpy as np
import pytest

import sfs_oracle as O
from helpers import close

pytestmark = pytest.mark.gpu


def rand_dict(rng):
    chroms = [f"c{int(x)}" for x in rng.choice(np.arange(1, 30), size=int(rng.integers(1, 5)), replace=False)]
    n1, n2 = int(rng.integers(1, 9)), int(rng.integers(1, 9))
    L = int(rng.choice([600, 4000, 30000]))
    nsnp = int(rng.choice([3, 40, 250]))
    miss = float(rng.choice([0.0, 0.1, 0.45]))
    p_missing_pop = float(rng.choice([0.0, 0.0, 0.2]))
    items = []
    for c in chroms:
        ps = rng.choice(np.arange(0 if rng.random() < 0.3 else 1, L), size=min(nsnp, L - 1), replace=False)
        for p in ps:
            calls = {}
            f = rng.random() ** 3 if rng.random() > 0.25 else 1 - rng.random() ** 3
            for pop, n in (("A", n1), ("B", n2)):
                if rng.random() < p_missing_pop:
                    continue
                called = 2 * n - 2 * rng.binomial(n, miss)
                alt = int(rng.binomial(called, min(max(f + rng.normal(0, 0.1), 0), 1)))
                calls[pop] = (int(called - alt), alt)
            items.append((f"{c}-{int(p)}", {"segregating": ("A", "C"), "context": "-A-", "calls": calls, "annotation": str(rng.choice(["x", "y"]))}))
    order = rng.permutation(len(items))
    d = {items[i][0]: items[i][1] for i in order}
    ctor = dict(pop1="A", pop2="B", pop1_size=n1, pop2_size=n2)
    r = rng.random()
    if r < 0.15:
        ctor["variant_type"] = "x"
    elif r < 0.3:
        ctor["start_position"], ctor["end_position"] = int(L * 0.2), int(L * 0.8)
    elif r < 0.4:
        ctor["fold"] = False
    W = int(rng.choice([max(L // 40, 1), max(L // 6, 1), L * 2]))
    N = int(rng.choice([2, 7, 30]))
    return d, ctor, chroms, W, N


def outcome(fn):
    try:
        return "ok", fn()
    except Exception as e:  # noqa: BLE001
        return "raises", type(e).__name__


def same(a, b, what):
    assert a[0] == b[0], (what, a[0], b[0], a[1] if a[0] == "raises" else "", b[1] if b[0] == "raises" else "")
    if a[0] == "raises":
        assert a[1] == b[1], (what, a[1], b[1])
        return
    ra, rb = a[1], b[1]
    assert list(ra) == list(rb), (what, list(ra)[:4], list(rb)[:4])
    for k in ra:
        assert list(ra[k]) == list(rb[k]), (what, k)
        for f in ra[k]:
            assert close(ra[k][f], rb[k][f]), (what, k, f, ra[k][f], rb[k][f])


import os
_SEEDS = range(int(os.environ.get("TDSFS_DIFF_SEED0", "0")), int(os.environ.get("TDSFS_DIFF_SEED1", "40")))


@pytest.mark.parametrize("seed", _SEEDS)
def test_random_dict_all_scanners(seed):
    import twoDSFS_class as K
    import sims_scan as S
    rng = np.random.default_rng(1000 + seed)
    d, ctor, chroms, W, N = rand_dict(rng)
    p = O.Panel(**ctor)
    mk = lambda: K.LikelihoodInference_jointSFS("x", "y", **ctor)  # noqa: E731
    # spectra
    raw2 = outcome(lambda: O.calculate_2d_sfs(d, p.pop1, p.pop2, p.n1, p.n2, p.start, p.end, p.vt, p.fold))
    got2 = outcome(lambda: mk().calculate_2d_sfs(d))
    assert raw2[0] == got2[0] == "ok" and raw2[1] == got2[1]
    for pop, n in (("A", p.n1), ("B", p.n2)):
        assert O.calculate_1d_sfs(d, pop, n, p.start, p.end, p.vt) == mk().calculate_1d_sfs(d, pop, n, p.start, p.end, p.vt)
    f1, f2 = p.s1(d, 1), p.s1(d, 2)
    same(outcome(lambda: mk().combined_scan(d, W)), outcome(lambda: O.combined_scan(p, d, W)), "combined_scan")
    same(outcome(lambda: mk().scan_perChr_bySNPs(d, N)), outcome(lambda: O.scan_perChr_bySNPs(p, d, N)), "scan_perChr_bySNPs")
    same(outcome(lambda: mk().scan_chooseChr(d, W, chroms[0])), outcome(lambda: O.scan_chooseChr(p, d, W, chroms[0])), "scan_chooseChr")
    same(outcome(lambda: mk().scan_chooseChr_bySNPs(d, N, chroms[-1])), outcome(lambda: O.scan_chooseChr_bySNPs(p, d, N, chroms[-1])), "scan_chooseChr_bySNPs")
    same(outcome(lambda: mk().scan_precomputed_BG(d, W, raw2[1], f1, f2)), outcome(lambda: O.scan_precomputed_BG(p, d, W, raw2[1], f1, f2)), "scan_precomputed_BG")
    same(outcome(lambda: mk().T2D_scan(d, raw2[1], W)), outcome(lambda: O.T2D_scan(p, d, raw2[1], W)), "T2D_scan")
    same(outcome(lambda: mk().T1D_scan(d, f1, W, "A", p.n1)), outcome(lambda: O.T1D_scan(p, d, f1, W, "A", p.n1)), "T1D_scan")
    same(outcome(lambda: mk().sims_process_window(d, W, raw2[1], f1, f2)), outcome(lambda: O.sims_process_window_class(p, d, W, raw2[1], f1, f2)), "sims_process_window")
    u1 = O.calculate_1d_sfs(d, "A", p.n1, None, None, None)
    u2 = O.calculate_1d_sfs(d, "B", p.n2, None, None, None)
    same(outcome(lambda: S.process_window(d, raw2[1], u1, u2, W, "A", "B", p.n1, p.n2, None, None, None)),
         outcome(lambda: O.sims_process_window(d, raw2[1], u1, u2, W, "A", "B", p.n1, p.n2, None, None, None)), "sims.process_window")
