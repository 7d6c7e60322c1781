"""Particle sharding across the GPUs of one box.

Test particles never interact, so the path shards by particle with no collective on the data
path: every rank holds a full ephemeris copy and a slice of the population.  The only
communication is plumbing: a barrier around timed regions and a max / sum of a few scalars.

`weak`   every rank integrates its own n_per_gpu particles (population seeded per rank)
`strong` one population of n_total particles, cut into contiguous slices
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def local_population(generator, n: int, seed: int, world: int, rank: int, scaling: str = "weak"):
    """Population slice of this rank.  generator(n, seed=...) -> array [n][6]."""
    if scaling == "weak":
        return generator(n, seed=seed + rank)
    if scaling == "strong":
        lo, hi = shard_bounds(n, world, rank)
        return generator(n, seed=seed)[lo:hi]
    raise ValueError("scaling must be 'weak' or 'strong'")


def reduce_max_sum(dist, maxima, sums, device=None):
    """All-reduce: element-wise MAX of `maxima`, SUM of `sums` over ranks (lists of floats).

    `dist` is torch.distributed (already initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(maxima), list(sums)
    import torch
    tm = torch.tensor(list(maxima), dtype=torch.float64, device=device)
    ts = torch.tensor(list(sums), dtype=torch.float64, device=device)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    return [float(x) for x in tm.tolist()], [float(x) for x in ts.tolist()]


def gather_states(dist, local_state: np.ndarray, world: int, rank: int):
    """Host gather of per-rank outputs on rank 0 (the only data movement between ranks)."""
    if dist is None or world == 1:
        return local_state
    out = [None] * world if rank == 0 else None
    dist.gather_object(local_state, out, dst=0)
    return np.concatenate(out, axis=0) if rank == 0 else None
