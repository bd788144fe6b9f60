"""N>1 path on CPU: world_size-2 gloo (the bench's barrier / max-over-ranks / stream sharding)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from object_detection_tracking_b200 import replicas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dist.barrier()
    mine = replicas.streams_for_rank(8, rank, world)
    t = replicas.max_over_ranks([1.0 + rank, 5.0 - rank])
    q.put((rank, mine, t, replicas.aggregate_fps(80, t[0], world)))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_sharding_and_max_time():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]       # disjoint, covering
    for r in res:
        assert r[2] == [2.0, 5.0]                                         # max over ranks
        assert r[3] == 80 * 2 / 2.0


def test_single_process_passthrough():
    assert replicas.max_over_ranks([3.0]) == [3.0]
    assert replicas.streams_for_rank(3, 0, 1) == [0, 1, 2]
