"""world_size-2 gloo test (CPU) of the N>1 host logic bench.py uses: per-rank frame arcs are disjoint and cover what
a single rank would render over the same number of frames; the step time is the max over ranks; tile ownership of the
screen-tile split partitions the tiles."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from softrast_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = sharding.frames_for_rank(step=3, frames_per_step=64, rank=rank, world=world, path_frames=1024)
    gathered = [None] * world
    dist.all_gather_object(gathered, frames.tolist())
    ms = sharding.reduce_max_ms(10.0 + 5.0 * rank, dist)
    tiles = sharding.tiles_for_rank(510, rank, world)
    tg = [None] * world
    dist.all_gather_object(tg, tiles.tolist())
    if rank == 0:
        out.put((gathered, ms, tg))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, ms, tg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = (np.array(g) for g in gathered)
    assert a.size == b.size == 64
    assert np.intersect1d(a, b).size == 0, "ranks must render different frames"
    assert np.array_equal(b, (a + 512) % 1024), "rank 1 walks the opposite arc of the closed camera path"
    assert ms == 15.0, "step time is the max over ranks"
    assert sorted(tg[0] + tg[1]) == list(range(510)) and not set(tg[0]) & set(tg[1])
