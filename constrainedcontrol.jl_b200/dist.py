"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The hot path shards by *systems*: every rank owns a contiguous slice of the batch and no collective runs on the step
path.  The only exchange is the all-gather of gain tables synthesised on different ranks when every rank needs all of
them (TrackingLQR / per-system LQR gains) -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """contiguous split of `total` systems: returns (lo, hi) of this rank (sizes differ by at most one)"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_gains(K_local: np.ndarray, total: int, device=None) -> np.ndarray:
    """K_local [S_local, ...] on every rank -> [total, ...] on every rank (order = rank order = batch order)"""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return K_local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    per = int(np.prod(K_local.shape[1:]))
    smax = max(hi - lo for lo, hi in sizes)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    buf = torch.zeros((smax, per), dtype=torch.float64, device=dev)
    buf[: K_local.shape[0]] = torch.from_numpy(np.ascontiguousarray(K_local).reshape(K_local.shape[0], per)).to(dev)
    out = torch.empty((world, smax, per), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out.view(world * smax, per), buf) if dist.get_backend() == "nccl" else dist.all_gather(list(out.unbind(0)), buf)
    parts = [out[r, : hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)]
    return np.concatenate(parts, axis=0).reshape((total,) + K_local.shape[1:])


def sharded_synthesis(synth_fn, total: int, device=None):
    """run `synth_fn(lo, hi) -> K [hi-lo, ...]` on this rank's shard and all-gather the gain tables"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        lo, hi = shard_range(total, dist.get_rank(), dist.get_world_size())
    else:
        lo, hi = 0, total
    return allgather_gains(synth_fn(lo, hi), total, device)
