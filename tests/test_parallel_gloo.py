"""World-size-2 gloo test of the multi-GPU layer (CPU): batch sharding + variable-length map all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gradslam_b200 as gs
from gradslam_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        sizes = [3 + rank, 7 - 2 * rank]  # ragged, different per rank
        mk = lambda c: [torch.rand(n, c, generator=g) for n in sizes]
        pc = gs.Pointclouds(mk(3), mk(3), mk(3), mk(1))
        keep = [p.clone() for p in pc.points_list]
        allpc = parallel.gather_maps(pc)
        counts = allpc.num_points_per_pointcloud.tolist()
        ok = counts == [3, 7, 4, 5] and len(allpc) == 4
        ok = ok and torch.equal(allpc.points_list[2 * rank], keep[0]) and torch.equal(allpc.points_list[2 * rank + 1], keep[1])
        pad = allpc.points_padded
        ok = ok and pad.shape == (4, 7, 3) and float(pad[0, 3:].abs().sum()) == 0.0
        # the same maps built inside a job-wide store: only the peer's rows move, the result is the store itself
        store = parallel.GatheredMaps(2, 9, "cpu")
        loc = store.reset()
        for b, n in enumerate(sizes):
            loc._geo[b, :n] = pc._geo[b, :n]
            loc._col[b, :n] = pc._col[b, :n]
        loc._set_counts(sizes)
        allpc2 = parallel.gather_maps(loc, into=store)
        ok = ok and allpc2 is store.all and allpc2.num_points_per_pointcloud.tolist() == counts
        for i in range(4):
            for name in ("points_list", "normals_list", "colors_list", "features_list"):
                ok = ok and torch.equal(getattr(allpc2, name)[i], getattr(allpc, name)[i])
        ok = ok and torch.equal(allpc2.points_padded, pad)
        lo, hi = parallel.shard_batch(5)
        ok = ok and (lo, hi) == ((0, 3) if rank == 0 else (3, 5))
        q.put((rank, bool(ok), counts))
    finally:
        dist.destroy_process_group()


def test_gather_maps_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    # both ranks see the same gathered sizes
    assert res[0][2] == res[1][2]


def test_shard_batch_covers_everything():
    for total in (1, 7, 8, 32, 33):
        for world in (1, 2, 4, 8):
            blocks = [parallel.shard_batch(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
