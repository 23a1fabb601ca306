"""2-GPU parity: a scan sharded by contiguous chromosome ranges with the background all-reduced (NCCL, and the library's
own peer-memory kernel) equals the single-GPU scan (integers bit for bit, fp64 to 1e-12).  Skipped on boxes with
fewer than two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _panel():
    from tdsfs_pack import pack_codes
    rng = np.random.default_rng(17)
    n1, n2, sizes = 40, 30, [3000, 5000, 2000, 4000, 3500, 2500]
    S = sum(sizes)
    f = np.exp(rng.uniform(np.log(0.005), np.log(0.995), size=S))

    def codes(ns):
        p = f[:, None]
        a = (rng.random((S, ns)) < p).astype(np.uint8) + (rng.random((S, ns)) < p).astype(np.uint8)
        c = np.where(a == 2, 3, a).astype(np.uint8)
        c[rng.random((S, ns)) < 0.03] = 2
        return c

    c1, c2 = codes(n1), codes(n2)
    pos = np.concatenate([np.sort(rng.choice(np.arange(1, 400000), size=s, replace=False)) for s in sizes]).astype(np.int32)
    return n1, n2, sizes, c1, c2, pos, pack_codes


def _worker(rank, world, port, out, peer=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.join(ROOT, "2dsfs-scan_b200"))
    import torch
    import torch.distributed as dist
    import tdsfs_capi as T
    from tdsfs_dist import shard_chromosomes, sharded_scan_bp, peer_setup, peer_teardown
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n1, n2, sizes, c1, c2, pos, pack_codes = _panel()
    off = np.concatenate([[0], np.cumsum(sizes)])
    lo, hi = shard_chromosomes(sizes, world)[rank]
    a, b = off[lo], off[hi]
    G, w1, w2 = pack_codes(c1[a:b], c2[a:b])
    h = T.Handle(rank)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, int(b - a), w1, w2, n1, n2, pos[a:b], off[lo:hi + 1] - off[lo])
    if peer:
        h.background(T.BG_GENOME)  # the histogram must exist before it can be exported
        assert peer_setup(h), "CUDA IPC mapping of the peers' histograms failed"
        for _ in range(3):         # repeated scans: barrier epochs advance, the histogram is re-zeroed every time
            res = sharded_scan_bp(h, 20000, T.BG_GENOME, torch.device("cuda", rank), chrom_base=lo, peer=True)
        h._check(h._L.tdsfs_check(h._h))
        peer_teardown(h)
    else:
        res = sharded_scan_bp(h, 20000, T.BG_GENOME, torch.device("cuda", rank), chrom_base=lo)
    if rank == 0:
        out["res"] = {k: v.tolist() for k, v in res.items()}
    dist.destroy_process_group()


@pytest.mark.parametrize("peer", [False, True], ids=["nccl", "peer-memory"])
def test_sharded_scan_equals_single_gpu(peer):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import tdsfs_capi as T
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out, peer), nprocs=2, join=True)
    n1, n2, sizes, c1, c2, pos, pack_codes = _panel()
    off = np.concatenate([[0], np.cumsum(sizes)])
    G, w1, w2 = pack_codes(c1, c2)
    h = T.Handle(0)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, len(pos), w1, w2, n1, n2, pos, off)
    single = h.run_bp(T.BG_GENOME, 20000)
    multi = {k: np.array(v) for k, v in out["res"].items()}
    for k, v in single.items():
        if v.dtype == np.float64:
            assert np.allclose(multi[k], v, rtol=1e-12, atol=1e-12, equal_nan=True), k
        else:
            assert np.array_equal(multi[k], v), k
