"""2-GPU parity: scans sharded by make_shard_plan (contiguous row ranges, split inside a chromosome on a window boundary) with
the background exchanged by NCCL or by the library's own peer-memory kernel equal the single-GPU scans (integers bit for
bit, fp64 to 1e-10).  Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`)."""
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
    n1, n2, sizes = 40, 30, [3000, 9000, 2000, 2500, 1500, 2000]  # the middle of the genome lies inside chromosome 1
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
    from tdsfs_dist import make_shard_plan, local_offsets, sharded_scan, peer_setup, peer_teardown
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n1, n2, sizes, c1, c2, pos, pack_codes = _panel()
    off = np.concatenate([[0], np.cumsum(sizes)])
    dev = torch.device("cuda", rank)
    results = {}
    h = T.Handle(rank)
    h.set_panel(n1, n2, True)
    loaded = None
    for name, size, snp, mode, bgc in CASES:
        plan = make_shard_plan(pos, off, world, **({"N": size} if snp else {"W": size}))
        pieces = plan[rank]
        key = tuple((p.lo, p.hi) for p in pieces)
        if key != loaded:  # rows of this rank's pieces, each piece one local chromosome (a split chromosome is on both ranks)
            rows = np.concatenate([np.arange(p.lo, p.hi) for p in pieces]) if pieces else np.zeros(0, np.int64)
            G, w1, w2 = pack_codes(c1[rows], c2[rows])
            h.load_genotypes(G, len(rows), w1, w2, n1, n2, pos[rows].astype(np.int32), local_offsets(pieces))
            loaded = key
        use_peer = peer and mode in (T.BG_GENOME, T.BG_CHROM)
        if use_peer:
            h.background(mode, -1 if mode == T.BG_CHROM else 0)  # the histogram must exist before it can be exported
            assert peer_setup(h), "CUDA IPC mapping of the peers' histograms failed"
        for _ in range(3 if use_peer else 1):  # repeated scans: barrier epochs advance, the histogram is re-zeroed every time
            res = sharded_scan(h, pieces, pos, size, mode, dev, snp_mode=snp, bg_chrom=bgc, plan=plan, rank=rank, peer=use_peer)
        h.check()
        if use_peer:
            peer_teardown(h)
        results[name] = {k: v.tolist() for k, v in res.items()}
    if peer:
        # tdsfs_step_bp on device-resident shards: the exchange runs in the count kernel's tail (after a grid barrier), the whole
        # pass is captured and replayed as a CUDA graph; every replay must leave the results of the call-by-call pass
        plan = make_shard_plan(pos, off, world, W=20000)
        pieces = plan[rank]
        rows = np.concatenate([np.arange(p.lo, p.hi) for p in pieces])
        G, w1, w2 = pack_codes(c1[rows], c2[rows])
        Gd, pd = torch.from_numpy(G.view(np.int32)).to(dev), torch.from_numpy(pos[rows].astype(np.int32)).to(dev)
        h.load_genotypes(Gd, len(rows), w1, w2, n1, n2, pd, local_offsets(pieces))
        h.background(T.BG_GENOME)
        assert peer_setup(h)
        h.plan(20000)
        h.background(T.BG_GENOME)
        h.peer_reduce_finalize()
        ref = h.scan(20000)
        h.set_sync(False)
        ok = True
        for i in range(5):
            h.step_bp(T.BG_GENOME, 20000)
            h.check()
            got = h.fetch_results(len(ref["start"]))
            for k, v in ref.items():
                same = np.allclose(got[k], v, rtol=1e-10, atol=1e-10, equal_nan=True) if v.dtype == np.float64 else np.array_equal(got[k], v)
                ok = ok and bool(same)
        h.set_sync(True)
        peer_teardown(h)
        out[f"step_ok_{rank}"] = ok
    if rank == 0:
        out["res"] = results
    dist.destroy_process_group()


# name, window size, fixed-SNP?, background mode (tdsfs_capi constants), GLOBAL background chromosome
CASES = [("bp_genome", 20000, False, 2, None), ("snp_genome", 300, True, 2, None), ("bp_chrom", 20000, False, 3, 1),
         ("bp_perchrom", 20000, False, 1, None), ("snp_perchrom", 177, True, 1, None)]


@pytest.mark.parametrize("peer", [False, True], ids=["nccl", "peer-memory"])
def test_sharded_scan_equals_single_gpu(peer):
    """2 GPUs, rows split by make_shard_plan (the 9000-SNP chromosome is split on a window boundary): fixed-bp and fixed-SNP
    windows, genome-wide / single-chromosome (global index) / per-chromosome backgrounds; gathered == single GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import tdsfs_capi as T
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out, peer), nprocs=2, join=True)
    if peer:
        assert out["step_ok_0"] and out["step_ok_1"], "tdsfs_step_bp (exchange in the count kernel's tail, graph replay) differs from the call-by-call pass"
    n1, n2, sizes, c1, c2, pos, pack_codes = _panel()
    off = np.concatenate([[0], np.cumsum(sizes)])
    G, w1, w2 = pack_codes(c1, c2)
    h = T.Handle(0)
    h.set_panel(n1, n2, True)
    h.load_genotypes(G, len(pos), w1, w2, n1, n2, pos, off)
    for name, size, snp, mode, bgc in CASES:
        h.plan(size, snp_mode=snp)
        h.background(mode, bgc if bgc is not None else 0)
        h.finalize_background()
        single = h.scan(size, snp_mode=snp)
        multi = {k: np.array(v) for k, v in out["res"][name].items()}
        for k, v in single.items():
            if v.dtype == np.float64:
                assert np.allclose(multi[k], v, rtol=1e-10, atol=1e-10, equal_nan=True), (name, k)
            else:
                assert np.array_equal(multi[k], v), (name, k)
    h.close()
