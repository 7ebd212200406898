"""Frame farm host logic on CPU: world_size 2, gloo. Each rank 'renders' its views with the oracle
(standing in for a GPU), frames are gathered to rank 0 and compared with a single-process render."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, num_views, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from edxraster_b200 import farm, scenes
    from oracle import orc
    sc = scenes.config4(width=160, height=96, quads_x=60, quads_z=48)
    views = scenes.config5_views(sc, num_views)
    mine = farm.views_of_rank(num_views, world, rank)
    o = orc.Oracle(sc.width, sc.height, 1)
    o.set_shader(1)
    frames = []
    for v in mine:
        o.set_transform(*views[v])
        o.render(sc.vertices, sc.indices)
        frames.append(torch.from_numpy(o.color()))
    local = torch.stack(frames) if frames else torch.zeros((0, sc.height, sc.width, 4), dtype=torch.uint8)
    out = farm.gather_frames(local, num_views, dst=0)
    if rank == 0:
        ref = []
        for v in range(num_views):
            o.set_transform(*views[v])
            o.render(sc.vertices, sc.indices)
            ref.append(o.color())
        q.put(bool((out.numpy() == np.stack(ref)).all()) and out.shape[0] == num_views)
    dist.barrier()
    dist.destroy_process_group()


def test_views_are_dealt_round_robin():
    sys.path.insert(0, ROOT)
    from edxraster_b200 import farm
    assert farm.views_of_rank(256, 8, 3) == list(range(3, 256, 8))
    got = sorted(sum((farm.views_of_rank(5, 2, r) for r in range(2)), []))
    assert got == [0, 1, 2, 3, 4]


def test_gather_to_rank0_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)
