"""CPU-only: stream sharding across ranks (gloo, world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kvazzup_b200 import sharding


def test_partition_is_disjoint_and_complete():
    for n in (0, 1, 7, 8, 240, 256):
        for world in (1, 2, 4, 8):
            parts = [sharding.streams_of_rank(n, r, world) for r in range(world)]
            flat = sorted(s for p in parts for s in p)
            assert flat == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
            assert all(sharding.rank_of_stream(s, world) == r for r, p in enumerate(parts) for s in p)
    with pytest.raises(ValueError):
        sharding.streams_of_rank(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.streams_of_rank(9, rank, world)
    frames, secs = sharding.gather_totals(len(mine) * 90, 1.0 + rank)
    q.put((rank, mine, frames, secs))
    dist.destroy_process_group()


def test_two_ranks_agree_on_totals_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7]
    assert all(r[2] == 9 * 90 and r[3] == 2.0 for r in res)       # frames summed, time = max over ranks
