"""Pose sharding across GPUs: every (complex, sample) pose is an independent unit, so ranks take
contiguous pose ranges balanced by estimated work, the weights are broadcast once and the final poses gathered once.
There is no communication inside a diffusion step.  Backend: ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU
tests)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of ``n`` items for ``rank`` (the first ``n % world`` ranks get one more)."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_poses(samples_per_complex: Sequence[int], cost_per_pose: Sequence[float], rank: int, world: int) -> List[Tuple[int, int, int]]:
    """Split the flattened (complex, sample) list into ``world`` contiguous ranges of ~equal estimated cost
    (e.g. N_l * N_r per pose); returns [(complex index, first sample, stop sample)] of ``rank``."""
    total = sum(s * c for s, c in zip(samples_per_complex, cost_per_pose))
    lo, hi = total * rank / world, total * (rank + 1) / world
    out, acc = [], 0.0
    for ci, (s, c) in enumerate(zip(samples_per_complex, cost_per_pose)):
        first = stop = None
        for k in range(s):
            mid = acc + 0.5 * c           # a pose belongs to the rank whose cost window holds its midpoint
            if lo <= mid < hi:
                first = k if first is None else first
                stop = k + 1
            acc += c
        if first is not None:
            out.append((ci, first, stop))
    return out


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """One broadcast per tensor of the (8 MB) parameter set; afterwards every rank packs its own weight blob."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def gather_poses(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """All-gather pose tensors ``[n_rank, ...]`` whose leading sizes ``counts`` are known on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
