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
