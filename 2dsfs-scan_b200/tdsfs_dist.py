"""Multi-GPU plumbing: one process per GPU (torchrun), windows sharded by contiguous chromosome ranges, and ONE
collective -- an all-reduce (sum) of the integer background histogram -- when the background spans ranks
(SURVEY.md section 8e).  Per-chromosome backgrounds with unsplit chromosomes need no collective at all.

torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests); the histogram lives in
libtdsfs' device memory and is all-reduced in place through a zero-copy tensor view.
"""
from __future__ import annotations

import numpy as np


def shard_chromosomes(sizes, world):
    """Contiguous chromosome ranges balanced by SNP count.  sizes: SNPs per chromosome (sorted chromosome order).
    Returns [(lo, hi)] per rank with lo <= hi, covering [0, C) in order (a rank may be empty when world > C)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    C = len(sizes)
    total = int(sizes.sum())
    bounds = [0]
    csum = np.concatenate([[0], np.cumsum(sizes)])
    for r in range(1, world):
        target = total * r / world
        # first boundary whose prefix is >= target, but keep at least monotone boundaries
        b = int(np.searchsorted(csum, target, side="left"))
        # choose the closer of b-1 / b to the target
        if b > 0 and abs(csum[b - 1] - target) <= abs(csum[min(b, C)] - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), C))
    bounds.append(C)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class _DevBuf:
    """__cuda_array_interface__ view of a raw device pointer (uint32 counts reinterpreted as int32 for torch)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def background_tensor(handle, device):
    """Zero-copy torch view of the handle's packed background histogram [group][2D | 1D pop1 | 1D pop2]."""
    import torch
    ptr, n, _ = handle.background_device()
    return torch.as_tensor(_DevBuf(ptr, n), device=device)


def allreduce_background(hist, group=None):
    """Sum the background histogram over ranks, in place.  Counts stay exact: int32 holds 2^31 - 1 SNPs per bin."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def peer_setup(handle, group=None):
    """Map every rank's background histogram into every other rank (CUDA IPC over NVLink) for
    handle.peer_allreduce_background().  Collective; the handle must hold its shard and tdsfs_background must have
    run once (single-group mode).  Returns False, leaving the NCCL path in use, when the mapping is not possible."""
    import torch.distributed as dist
    import tdsfs_capi as T
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    try:
        blob = handle.peer_export(rank, world)
    except T.TdsfsError:
        blob = None
    blobs = [None] * world
    dist.all_gather_object(blobs, blob, group=group)
    ok = all(b is not None for b in blobs)
    if ok:
        try:
            handle.peer_import(blobs)
        except T.TdsfsError:
            ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)  # also the host barrier: every rank has mapped before anyone signals
    if not all(flags):
        handle.peer_close()
        return False
    return True


def peer_teardown(handle, group=None):
    """Collective: unmap the peers' histograms before any rank frees or re-shapes its own."""
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    dist.barrier(group=group)
    handle.peer_close()
    dist.barrier(group=group)


def gather_results(local, chrom_base, group=None):
    """Concatenate per-rank result arrays in rank order (== the reference's sorted window order, because ranks own
    contiguous chromosome ranges).  local: dict name -> numpy array; chrom_base: global index of this rank's first chromosome."""
    import torch.distributed as dist
    local = dict(local)
    if "chrom" in local:
        local["chrom"] = local["chrom"] + np.int32(chrom_base)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    return {k: np.concatenate([p[k] for p in parts]) for k in local}


def sharded_scan_bp(handle, window_bp, bg_mode, device, group=None, chrom_base=0, peer=False):
    """background -> (all-reduce) -> finalize -> scan -> gather, for a handle that already holds this rank's shard.
    peer=True: the all-reduce is the library's own peer-memory kernel (after peer_setup) instead of NCCL."""
    import tdsfs_capi as T
    handle.plan(window_bp)  # window boundaries on a side stream: overlaps the count kernel and the all-reduce
    handle.background(bg_mode)
    if bg_mode in (T.BG_GENOME, T.BG_CHROM):
        if peer:
            handle.peer_allreduce_background()
        else:
            allreduce_background(background_tensor(handle, device), group)
    handle.finalize_background()
    return gather_results(handle.scan(window_bp), chrom_base, group)
