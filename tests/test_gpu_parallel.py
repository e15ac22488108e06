"""Two-GPU test of the map exchange: the three transports (copy-engine pulls through IPC
mappings, NCCL all-gather, NCCL send/recv) must deliver identical maps.  Skipped on boxes with one GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import gradslam_b200 as gs
    from gradslam_b200 import parallel
    from gradslam_b200.synthetic import make_sequence

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, L, H, W = 2, 3 + rank, 48, 64  # different map sizes per rank
        rgb, depth, K, poses = make_sequence(B, L, H, W, seed=40 + rank)
        slam = gs.PointFusion(odom="gt", device=dev)
        got = {}
        for mode in ("peer", "all_gather", "p2p"):
            os.environ["GSX_MAP_EXCHANGE"] = mode
            for rep in range(2):  # second round: cached IPC mappings, recycled stores
                pc, _ = slam(gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev)))
                allpc = parallel.gather_maps(pc)
                torch.cuda.synchronize(dev)
            counts = allpc.num_points_per_pointcloud.tolist()
            own = all(torch.equal(allpc.points_list[rank * B + b], pc.points_list[b]) and
                      torch.equal(allpc.colors_list[rank * B + b], pc.colors_list[b]) and
                      torch.equal(allpc.features_list[rank * B + b], pc.features_list[b]) for b in range(B))
            got[mode] = (counts, own, [t.cpu() for t in allpc.points_list], [t.cpu() for t in allpc.normals_list],
                         [t.cpu() for t in allpc.colors_list], allpc.points_padded.cpu())
        # job-wide store: the sequences are fused in place into this rank's block, only the peers' rows travel
        os.environ["GSX_MAP_EXCHANGE"] = "peer"
        P = H * W
        store = parallel.GatheredMaps(B, 5 * P, dev)  # room for the longer of the two ranks' sequences
        for rep in range(2):
            pc, _ = slam(gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev)), out=store.reset())
            assert pc is store.local
            allpc = parallel.gather_maps(pc, into=store)
            torch.cuda.synchronize(dev)
        assert allpc is store.all
        got["in_place"] = (allpc.num_points_per_pointcloud.tolist(), True, [t.cpu() for t in allpc.points_list],
                           [t.cpu() for t in allpc.normals_list], [t.cpu() for t in allpc.colors_list],
                           allpc.points_padded.cpu())
        ref = got["all_gather"]
        ok = True
        for mode in ("peer", "p2p", "in_place"):
            g = got[mode]
            ok = ok and g[0] == ref[0] and g[1] and ref[1]
            for k in (2, 3, 4):
                ok = ok and all(torch.equal(a, b) for a, b in zip(g[k], ref[k]))
            ok = ok and torch.equal(g[5], ref[5])  # zero padding of the ragged tails restored in every mode
        q.put((rank, bool(ok), ref[0]))
    finally:
        os.environ.pop("GSX_MAP_EXCHANGE", None)
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_map_exchange_transports_agree_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2]
