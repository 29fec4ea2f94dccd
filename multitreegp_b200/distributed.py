"""Population sharding over the GPUs of one box (one process per GPU, torch.distributed).

Mirrors the reference's only parallelism: a 1-D device mesh over the population axis,
``shard_map(in_specs=(P('i'), P(None)), out_specs=P('i'))`` (genetic_programming.py:255-262): rank r
evaluates the contiguous slice [r*P/n, (r+1)*P/n) of the flattened population, the data tuple is
replicated, and the fitness shards are concatenated in rank order -- here with ONE all-gather of P/n
floats per rank (NCCL over NVLink on GPUs, gloo in the CPU tests).  There is no exchange inside the
evaluation itself: individuals are independent.
"""
from __future__ import annotations

from typing import Callable

import numpy as np


def shard_bounds(P: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous slice of rank `rank`; like shard_map, the population must divide evenly."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    if P % world_size != 0:
        raise ValueError(f"population size {P} is not divisible by the number of devices {world_size} "
                         "(shard_map over the population axis requires it, genetic_programming.py:259-262)")
    n = P // world_size
    return rank * n, (rank + 1) * n


def evaluate_population_sharded(flat_populations, data, evaluate_shard: Callable, size_parsimony: float = 0.0,
                                group=None):
    """fitness f32[P] on every rank.  ``evaluate_shard(pop_slice, data, size_parsimony) -> fitness``
    runs on this rank's device (a CUDA torch tensor on GPUs).  Works with torch tensors (GPU path,
    NCCL) and numpy arrays (gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    P = flat_populations.shape[0]
    lo, hi = shard_bounds(P, world, rank)
    local = evaluate_shard(flat_populations[lo:hi], data, size_parsimony)
    if world == 1:
        return local
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local.contiguous()
    out = torch.empty((P,), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.numpy() if is_np else out
