"""The N>1 path on CPU: world_size-2 gloo processes exercise the sharding and the only
communication the path has (max / sum of scalars, host gather of outputs)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from assist_b200 import sharding
from assist_b200.synth import populations


def test_shard_bounds_cover_and_are_disjoint():
    for n, w in ((10, 3), (1000000, 8), (7, 8), (5, 1)):
        spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_strong_slices_reassemble_the_population():
    full = populations.neo_mba_mix(999, seed=7)
    parts = [sharding.local_population(populations.neo_mba_mix, 999, 7, 4, r, "strong") for r in range(4)]
    assert np.array_equal(sharding.interleave(parts), full)
    assert np.array_equal(parts[1], full[1::4])
    a = sharding.local_population(populations.neo_mba_mix, 100, 7, 2, 0, "weak")
    b = sharding.local_population(populations.neo_mba_mix, 100, 7, 2, 1, "weak")
    assert a.shape == b.shape == (100, 6) and not np.array_equal(a, b)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from assist_b200 import sharding as sh
    from assist_b200.synth import populations as pop
    local = sh.local_population(pop.main_belt, 101, 11, world, rank, "strong")
    # stand-in for the per-rank integration: a deterministic function of the local slice
    result = local * 2.0
    maxima, sums = sh.reduce_max_sum(dist, [1.0 + rank, 5.0 - rank], [float(local.shape[0]), 3.0])
    gathered = sh.gather_states(dist, result, world, rank, "strong")
    rows = sh.gather_rows(dist, [float(rank), float(local.shape[0])])
    if rank == 0:
        q.put((maxima, sums, gathered, rows))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_and_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    maxima, sums, gathered, rows = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert maxima == [2.0, 5.0] and sums == [101.0, 6.0]
    assert np.array_equal(gathered, populations.main_belt(101, seed=11) * 2.0)
    assert rows == [[0.0, 51.0], [1.0, 50.0]]
