"""Multi-GPU plumbing: robots are independent (each has its own World,
Locomotion.java:172), so the population is block-partitioned over ranks with no
per-step traffic; the only collective is the final gather of the fixed-size
``vsr_result`` records (Outcome summaries)."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_robots: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of robots owned by ``rank`` (sizes differ by <= 1)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("invalid rank / world_size")
    base, rem = divmod(n_robots, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class _DevicePtr:
    """Minimal ``__cuda_array_interface__`` holder for a raw device pointer."""

    def __init__(self, ptr: int, n_doubles: int):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def device_records(batch, device):
    """The device-resident ``vsr_result[n]`` of a DeviceBatch as a torch float64 [n, 11] view
    (no copy): the buffer the final NCCL gather sends from."""
    import torch

    ptr, nbytes = batch.results_device()
    t = torch.as_tensor(_DevicePtr(ptr, nbytes // 8), device=device)
    return t.view(-1, 11)


def gather_results(local, n_robots: int, device=None):
    """all_gather the per-rank ``vsr_result`` records into batch order on every rank.

    ``local`` is a numpy structured array (host records) or a torch float64 [n, 11] tensor
    (``device_records``: NCCL then sends straight from the library's result buffer).
    Returns a numpy structured array of ``n_robots`` records.
    """
    import torch
    import torch.distributed as dist

    from .hmsrobots import RESULT_DTYPE

    ws = dist.get_world_size() if dist.is_initialized() else 1
    words = RESULT_DTYPE.itemsize // 8
    if isinstance(local, np.ndarray):
        t = torch.from_numpy(local.view(np.float64).reshape(-1, words).copy())
        if device is not None:
            t = t.to(device)
    else:
        t = local.view(torch.float64).reshape(-1, words)
    if ws == 1:
        return t.cpu().numpy().reshape(-1).view(RESULT_DTYPE).copy()
    sizes = [shard_range(n_robots, r, ws) for r in range(ws)]
    counts = [e - b for b, e in sizes]
    if min(counts) == max(counts):  # equal shards: one collective into one buffer, one D2H
        out = torch.empty((n_robots, words), dtype=torch.float64, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out.cpu().numpy().reshape(-1).view(RESULT_DTYPE).copy()
    maxn = max(counts)
    pad = torch.zeros((maxn, words), dtype=torch.float64, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    parts = [o[:c].cpu().numpy() for o, c in zip(out, counts)]
    return np.concatenate(parts).reshape(-1).view(RESULT_DTYPE).copy()
