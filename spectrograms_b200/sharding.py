"""Clip sharding across the GPUs of one box.

The path shards by clip with no exchange step: rank r of G owns a contiguous block of
``ceil(n_clips / G)`` clips, runs its own plan replica on its own device, and never talks to the others on the hot
path. ``gather_to_rank0`` is the optional, off-the-hot-path collection of results (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_clips: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open clip range [lo, hi) owned by ``rank``: contiguous blocks of ceil(n_clips/world_size)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("rank/world_size out of range")
    per = -(-int(n_clips) // world_size)
    lo = min(rank * per, n_clips)
    hi = min(lo + per, n_clips)
    return lo, hi


def bind_host_to_device(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node), so that
    the pinned host buffers a rank allocates afterwards are first-touched next to its own GPU's PCIe root. One process
    per GPU otherwise leaves placement to chance and the ranks' host<->device copies share one socket's memory and
    fabric. Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(int(device_index)))
        return True
    except Exception:
        return False


def gather_to_rank0(local, n_clips: int, group=None):
    """Optional result gather (torch.distributed; backend nccl on GPUs, gloo on CPU): rank 0 receives every rank's
    block, concatenated in clip order; other ranks get ``None``. ``local`` is a torch tensor (n_local, rows, frames).
    Uneven tails are handled by padding the last block to ceil(n_clips/world) clips and trimming after the gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-int(n_clips) // world)
    pad = per - local.shape[0]
    if pad > 0:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], dim=0)
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat(bufs, dim=0)[:n_clips]
