"""world_size-2 gloo test of the N>1 host logic (the kernels themselves need no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "open-groundingdino_b200"))
    from msda_b200 import dist_utils
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = dist_utils.shard_bounds(7, world, rank)
        owned = torch.zeros(7)
        owned[lo:hi] = 1
        dist.all_reduce(owned)
        t = dist_utils.max_over_ranks(1.0 + rank)          # rank 1 is the slow one
        total = dist_utils.aggregate_throughput(units_this_rank=float(hi - lo), seconds_this_rank=1.0 + rank)
        out.put((rank, (lo, hi), owned.tolist(), t, total))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_partition_and_timing_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res[0][1] == (0, 4) and res[1][1] == (4, 7)
    for r in res:
        assert r[2] == [1.0] * 7          # every batch element owned exactly once
        assert r[3] == 2.0                # max over ranks of the elapsed time
        assert r[4] == pytest.approx(7 / 2.0)


def test_shard_bounds_cover_everything():
    from msda_b200.dist_utils import shard_bounds
    for n in (0, 1, 4, 7, 32):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
