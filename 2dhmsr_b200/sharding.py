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


def gather_results(local, n_robots: int, device=None):
    """all_gather the per-rank ``vsr_result`` records into batch order on every rank.

    ``local`` is either a numpy structured array (host records) or a torch uint8/float64 tensor
    viewing the device-resident records (``vsr_batch_results_device``), in which case NCCL sends
    straight from that buffer.  Returns a numpy structured array of ``n_robots`` records.
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
    maxn = max(e - b for b, e in sizes)
    pad = torch.zeros((maxn, words), dtype=torch.float64, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    parts = [o[: e - b].cpu().numpy() for o, (b, e) in zip(out, sizes)]
    return np.concatenate(parts).reshape(-1).view(RESULT_DTYPE).copy()
