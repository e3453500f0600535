"""N>1 host logic on CPU: world_size-2 gloo processes exercise request sharding, the image gather and the max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seedx_b200 import dist as sd
    mine = sd.shard_requests(7)
    img = torch.full((2, 4, 4, 3), 10 * rank + 1, dtype=torch.uint8)
    allimg = sd.gather_images(img)
    t = sd.max_over_ranks(100.0 + rank, "cpu")
    sd.barrier()
    q.put((rank, mine, allimg[:, 0, 0, 0, 0].tolist(), t))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    assert sorted(res[0][1] + res[1][1]) == list(range(7))
    for r in res:
        assert r[2] == [1, 11] and r[3] == 101.0


def test_single_process_defaults():
    from seedx_b200 import dist as sd
    assert sd.world() == 1 and sd.rank() == 0
    assert sd.shard_requests(3) == [0, 1, 2]
    assert sd.gather_images(torch.zeros((1, 2, 2, 3), dtype=torch.uint8)).shape == (1, 1, 2, 2, 3)
    assert sd.max_over_ranks(3.5, "cpu") == 3.5
