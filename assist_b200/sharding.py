"""Particle sharding across the GPUs of one box.

Test particles never interact, so the path shards by particle with no collective on the data
path: every rank holds a full ephemeris copy and a slice of the population.  The only
communication is plumbing: a barrier around timed regions and a max / sum of a few scalars.

`weak`   every rank integrates its own n_per_gpu particles (population seeded per rank)
`strong` one population of n_total particles dealt out round-robin: rank r integrates particles r, r + N, r + 2N ...
         (SURVEY section 8d, C3).  A population is usually ORDERED by kind (the C3 bench population holds its NEOs,
         six times the steps of a main-belt object, first): contiguous slices would give one GPU all the long
         systems, the round-robin deal gives every GPU the same mix.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def strong_indices(n_total: int, world: int, rank: int):
    """Indices of rank `rank` in the round-robin deal of n_total particles."""
    return np.arange(int(rank), int(n_total), int(world))


def local_population(generator, n: int, seed: int, world: int, rank: int, scaling: str = "weak"):
    """Population slice of this rank.  generator(n, seed=...) -> array [n][6]."""
    if scaling == "weak":
        return generator(n, seed=seed + rank)
    if scaling == "strong":
        return generator(n, seed=seed)[strong_indices(n, world, rank)]
    raise ValueError("scaling must be 'weak' or 'strong'")


def interleave(parts):
    """Inverse of the round-robin deal: parts[r] holds particles r, r + N, ... -> the population in its own order."""
    world = len(parts)
    n = sum(p.shape[0] for p in parts)
    out = np.empty((n,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        out[r::world] = p
    return out


def gather_rows(dist, row, device=None):
    """Every rank contributes a vector of floats; returns the [world][k] table on every rank (plumbing only:
    per-rank timings and counts of the bench line)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(row)]
    import torch
    mine = torch.tensor(list(row), dtype=torch.float64, device=device)
    table = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(table, mine)
    return [[float(x) for x in t.tolist()] for t in table]


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


def gather_states(dist, local_state: np.ndarray, world: int, rank: int, scaling: str = "weak"):
    """Host gather of per-rank outputs on rank 0 (the only data movement between ranks); a strong-scaled run gets
    its population back in the population's own order."""
    if dist is None or world == 1:
        return local_state
    out = [None] * world if rank == 0 else None
    dist.gather_object(local_state, out, dst=0)
    if rank != 0:
        return None
    return interleave(out) if scaling == "strong" else np.concatenate(out, axis=0)
