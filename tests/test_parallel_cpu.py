"""World-size-2 gloo test of the sharding helpers (the N > 1 path of bench.py / the room driver), CPU only."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from spsg_b200 import parallel as P
    r, w = P.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    chunks = list(P.shard_range(13, r, w))
    windows = list(P.shard_round_robin(13, r, w))
    gathered = [None] * w
    dist.all_gather_object(gathered, (chunks, windows))
    slowest = P.max_over_ranks(10.0 + 5.0 * rank)
    total = P.sum_over_ranks(len(chunks) * 5 * 81920)
    P.barrier()
    if rank == 0:
        out.put((gathered, slowest, total))
    dist.destroy_process_group()


def test_sharding_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, slowest, total = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    chunks = [g[0] for g in gathered]
    windows = [g[1] for g in gathered]
    assert sorted(chunks[0] + chunks[1]) == list(range(13)) and not set(chunks[0]) & set(chunks[1])
    assert chunks[0] == list(range(0, 7)) and chunks[1] == list(range(7, 13))   # contiguous, sizes differ by <= 1
    assert windows[0] == list(range(0, 13, 2)) and windows[1] == list(range(1, 13, 2))
    assert slowest == 15.0                      # max over ranks, the bench's timing rule
    assert total == 13 * 5 * 81920              # whole-job rays


def test_sharding_edge_cases():
    from spsg_b200 import parallel as P
    assert list(P.shard_range(0, 0, 4)) == []
    assert [len(P.shard_range(3, r, 8)) for r in range(8)] == [1, 1, 1, 0, 0, 0, 0, 0]
    assert sum(len(P.shard_range(1000003, r, 8)) for r in range(8)) == 1000003
    assert P.max_over_ranks(3.5) == 3.5 and P.sum_over_ranks(7) == 7.0   # single process: identity
    try:
        P.shard_range(4, 2, 2)
        assert False
    except ValueError:
        pass
