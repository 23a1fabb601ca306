"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chromosome sharding, the background all-reduce and the
rank-ordered result gather.  The device kernels are not involved (no GPU here); the collective runs on CPU tensors."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sfs_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_chromosomes_properties():
    from tdsfs_dist import shard_chromosomes
    rng = np.random.default_rng(0)
    for _ in range(200):
        C = int(rng.integers(1, 40))
        world = int(rng.integers(1, 9))
        sizes = rng.integers(0, 1000, size=C)
        sh = shard_chromosomes(sizes, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == C
        for (a, b), (c, d) in zip(sh, sh[1:]):
            assert a <= b == c <= d
    # balanced when chromosomes are equal
    sh = shard_chromosomes([100] * 32, 8)
    assert [b - a for a, b in sh] == [4] * 8
    sh = shard_chromosomes([100] * 32, 3)
    assert sorted(b - a for a, b in sh) == [10, 11, 11]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (os.path.join(ROOT, "2dsfs-scan_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tdsfs_dist import allreduce_background, gather_results, shard_chromosomes
    rng = np.random.default_rng(5)
    n1, n2, C = 6, 5, 5
    sizes = [300, 500, 200, 400, 350]
    S = sum(sizes)
    cnt = np.zeros((S, 4), dtype=np.int64)
    for p, n in ((0, n1), (1, n2)):
        called = 2 * n - 2 * rng.binomial(n, 0.1, S)
        alt = rng.binomial(called, rng.random(S) ** 2)
        cnt[:, 2 * p], cnt[:, 2 * p + 1] = called - alt, alt
    off = np.concatenate([[0], np.cumsum(sizes)])
    pos = np.concatenate([np.sort(rng.choice(np.arange(1, 50000), size=s, replace=False)) for s in sizes])
    lo, hi = shard_chromosomes(sizes, world)[rank]
    a, b = off[lo], off[hi]
    # local integer background of this rank's chromosomes (what the count kernel produces per rank) ...
    h2, h1, h1b = O.dense_spectra(cnt[a:b], n1, n2)
    packed = torch.from_numpy(np.concatenate([h2.ravel(), h1, h1b]).astype(np.int32))
    # ... all-reduced in place == the genome-wide background
    allreduce_background(packed)
    g2, g1, g1b = O.dense_spectra(cnt, n1, n2)
    assert np.array_equal(packed.numpy(), np.concatenate([g2.ravel(), g1, g1b]))
    # local windows scored against the GLOBAL background, gathered in rank order == single-process scan
    full = O.scan_arrays(cnt, pos, off, n1, n2, W=5000, bg="genome")
    loc_off = off[lo:hi + 1] - off[lo]
    loc = {"chrom": [], "start": [], "snp_count": [], "T2D": []}
    for c, s, l, h in O.bp_window_ranges(pos[a:b], loc_off, 5000):
        w2, _, _ = O.dense_spectra(cnt[a:b][l:h], n1, n2)
        T, none = O.clr_dense(w2.ravel()[1:-1], g2.ravel()[1:-1])
        loc["chrom"].append(c); loc["start"].append(s); loc["snp_count"].append(h - l); loc["T2D"].append(T)
    loc = {k: np.array(v, dtype=np.int32 if k == "chrom" else None) for k, v in loc.items()}
    allres = gather_results(loc, lo)
    ok = (np.array_equal(allres["chrom"], full["chrom"]) and np.array_equal(allres["start"], full["start"])
          and np.array_equal(allres["snp_count"], full["snp_count"]) and np.allclose(allres["T2D"], full["T2D"], rtol=1e-12, equal_nan=True))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world2_allreduce_and_gather_match_single_process():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


# ------------------------------------------------------------------------------------------------ shard plan + sharded_scan
class OracleHandle:
    """CPU stand-in of tdsfs_capi.Handle for the host logic of tdsfs_dist: same call sequence, the numerics by the oracle.
    One entry per CANDIDATE window (empty ones flagged), exactly like the C ABI."""

    def __init__(self, cnt, pos, off, n1, n2):
        self.cnt, self.pos, self.off, self.n1, self.n2 = cnt, np.asarray(pos), np.asarray(off, dtype=np.int64), n1, n2
        self.C = len(self.off) - 1
        self.gstride = (2 * n1 + 1) * (2 * n2 + 1) + 2 * n1 + 1 + 2 * n2 + 1
        self.hist = None

    def plan(self, size, snp_mode=False):
        pass

    def _pack(self, lo, hi):
        h2, h1, h1b = O.dense_spectra(self.cnt[lo:hi], self.n1, self.n2)
        return np.concatenate([h2.ravel(), h1, h1b]).astype(np.int32)

    def background(self, mode, bg_chrom=0):
        self.mode = mode
        if mode == 1:  # per chromosome
            self.hist = np.zeros((max(self.C, 1), self.gstride), dtype=np.int32)
            for c in range(self.C):
                self.hist[c] = self._pack(self.off[c], self.off[c + 1])
        else:
            self.hist = np.zeros((1, self.gstride), dtype=np.int32)
            if mode == 2 and self.C:
                self.hist[0] = self._pack(0, self.off[-1])
            elif mode == 3 and bg_chrom >= 0:
                self.hist[0] = self._pack(self.off[bg_chrom], self.off[bg_chrom + 1])

    def background_tensor(self):
        return torch.from_numpy(self.hist.reshape(-1))

    def finalize_background(self):
        pass

    def check(self):
        pass

    def scan(self, size, snp_mode=False):
        n1, n2 = self.n1, self.n2
        nb2 = (2 * n1 + 1) * (2 * n2 + 1)
        rows = []
        for c in range(self.C):
            lo, hi = int(self.off[c]), int(self.off[c + 1])
            if hi == lo:
                continue
            g = self.hist[c if self.mode == 1 else 0]
            b2 = g[:nb2].astype(np.int64)
            p = self.pos[lo:hi]
            if snp_mode:
                wins = [(j, lo + j * size, lo + (j + 1) * size) for j in range((hi - lo) // size)]
            else:
                ncand = max(int(p[-1]) - 1, 0) // size + 1
                k = np.maximum(p - 1, 0) // size
                wins = [(kk, lo + int(np.searchsorted(k, kk, "left")), lo + int(np.searchsorted(k, kk, "right"))) for kk in range(ncand)]
            for kk, a, b in wins:
                if snp_mode:
                    start = int(self.pos[a]) if kk == 0 else int(self.pos[a - 1]) + 1
                    end = int(self.pos[b - 1])
                else:
                    start, end = 1 + kk * size, (kk + 1) * size
                if b == a:
                    rows.append((c, start, end, 0, 0.0, F_EMPTY))
                    continue
                w2, _, _ = O.dense_spectra(self.cnt[a:b], n1, n2)
                T, none = O.clr_dense(w2.ravel()[1:-1], b2[1:-1])
                rows.append((c, start, end, b - a, np.nan if none else T, 1 if none else 0))
        cols = list(zip(*rows)) if rows else [[]] * 6
        return dict(chrom=np.array(cols[0], dtype=np.int32), start=np.array(cols[1], dtype=np.int64), end=np.array(cols[2], dtype=np.int64),
                    snp_count=np.array(cols[3], dtype=np.int32), T2D=np.array(cols[4], dtype=np.float64), flags=np.array(cols[5], dtype=np.uint8))


F_EMPTY = 8


def _uneven_genome(seed=3):
    """ECB-like: many contigs of very uneven sizes (a few large chromosomes, a tail of small scaffolds)."""
    rng = np.random.default_rng(seed)
    n1, n2 = 4, 3
    sizes = [2600, 1900, 40, 7, 1200, 3, 900, 55, 12, 300, 1, 1, 25, 600]
    S = sum(sizes)
    cnt = np.zeros((S, 4), dtype=np.int64)
    for p, n in ((0, n1), (1, n2)):
        called = 2 * n - 2 * rng.binomial(n, 0.1, S)
        alt = rng.binomial(called, rng.random(S) ** 2)
        cnt[:, 2 * p], cnt[:, 2 * p + 1] = called - alt, alt
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pos = np.concatenate([np.sort(rng.choice(np.arange(0, 40 * s + 50), size=s, replace=False)) for s in sizes]).astype(np.int64)
    return n1, n2, cnt, pos, off


def test_shard_plan_properties():
    from tdsfs_dist import make_shard_plan
    n1, n2, cnt, pos, off = _uneven_genome()
    S = int(off[-1])
    for world in (1, 2, 3, 8, 13):
        for kw in (dict(W=700), dict(N=50), dict(W=100000)):
            plan = make_shard_plan(pos, off, world, **kw)
            assert len(plan) == world
            rows = [(p.lo, p.hi) for pieces in plan for p in pieces]
            assert rows[0][0] == 0 and rows[-1][1] == S and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))  # a partition, in order
            for pieces in plan:
                for p in pieces:
                    assert off[p.chrom] <= p.lo < p.hi <= off[p.chrom + 1] and p.first == (p.lo == off[p.chrom])
                    if not p.first:  # split inside a chromosome: only on a window boundary
                        if "N" in kw:
                            assert (p.lo - off[p.chrom]) % kw["N"] == 0
                        else:
                            W = kw["W"]
                            assert max(pos[p.lo] - 1, 0) // W > max(pos[p.lo - 1] - 1, 0) // W
            if world == 8 and "W" in kw and kw["W"] == 700:  # balance: the large contigs are split
                loads = [sum(p.hi - p.lo for p in pieces) for pieces in plan]
                assert max(loads) <= 1.35 * S / world, loads


def _sharded_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (os.path.join(ROOT, "2dsfs-scan_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tdsfs_dist import make_shard_plan, local_offsets, sharded_scan
    n1, n2, cnt, pos, off = _uneven_genome()
    ok = {}
    for name, size, snp, mode, bgc in (("bp_genome", 700, False, 2, None), ("snp_genome", 50, True, 2, None), ("bp_chrom", 700, False, 3, 4),
                                       ("bp_perchrom", 700, False, 1, None), ("snp_perchrom", 37, True, 1, None), ("snp_chrom", 50, True, 3, 0)):
        plan = make_shard_plan(pos, off, world, **({"N": size} if snp else {"W": size}))
        pieces = plan[rank]
        rows = np.concatenate([np.arange(p.lo, p.hi) for p in pieces]) if pieces else np.zeros(0, np.int64)
        h = OracleHandle(cnt[rows], pos[rows], local_offsets(pieces), n1, n2)
        got = sharded_scan(h, pieces, pos, size, mode, "cpu", snp_mode=snp, bg_chrom=bgc, plan=plan, rank=rank)
        # single process on everything
        full = OracleHandle(cnt, pos, off, n1, n2)
        full.background(mode, bgc if bgc is not None else 0)
        exp = full.scan(size, snp_mode=snp)
        same = all(np.array_equal(got[k], exp[k]) for k in ("chrom", "start", "end", "snp_count", "flags"))
        same = same and np.allclose(got["T2D"], exp["T2D"], rtol=1e-12, atol=0, equal_nan=True)
        ok[name] = bool(same)
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_scan_equals_single_process(world):
    """world 2 and 8 (gloo), uneven ECB-like contigs: fixed-bp / fixed-SNP windows with a genome-wide background, a
    single-chromosome background given by its GLOBAL index, and per-chromosome backgrounds with split chromosomes --
    gathered arrays equal the single-process arrays entry for entry (labels, snp_count, empty candidates, statistics)."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        assert all(out[r].values()), (r, dict(out[r]))
