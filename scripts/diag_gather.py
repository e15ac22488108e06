"""torchrun diagnostic: where does the multi-GPU step time go?  (compute only / blocking gather / pipelined gather)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
import torch
import torch.distributed as dist

import gradslam_b200 as gs
from gradslam_b200 import parallel
from gradslam_b200.synthetic import make_sequence

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, L, H, W = 8, 32, 480, 640
rgb, depth, K, poses = make_sequence(B, L, H, W, seed=rank)
frames = gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev))
slam = gs.PointFusion(odom="gt", device=dev)


def run(mode, steps=6):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    pending = None
    for _ in range(steps):
        pc, _p = slam(frames)
        if mode == "blocking":
            parallel.gather_maps(pc)
        elif mode == "pipelined":
            if pending is not None:
                parallel.gather_maps_end(pending, wait=False)
            pending = parallel.gather_maps_begin(pc)
    if pending is not None:
        parallel.gather_maps_end(pending, wait=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    wall = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-10s %.2f ms/step (device, max over ranks)  %.2f ms/step wall  alloc %.1f GB reserved %.1f GB" % (
            mode, t.item(), wall, torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9), flush=True)


if rank == 0:
    print("NCCL_MAX_CTAS =", os.environ.get("NCCL_MAX_CTAS"), flush=True)
for mode in ("compute", "pipelined", "pipelined", "pipelined", "compute"):
    run(mode)
# phases of one blocking gather
pc, _p = slam(frames)
torch.cuda.synchronize()
t0 = time.perf_counter()
h = parallel.gather_maps_begin(pc)
h.ready.synchronize()
t1 = time.perf_counter()
out = parallel.gather_maps_end(h, wait=True)
torch.cuda.synchronize()
t2 = time.perf_counter()
if rank == 0:
    print("phases: counts %.2f ms, data %.2f ms, nmax %d, gathered maps %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3,
                                                                              out.points_padded.shape[1], len(out)))
dist.destroy_process_group()
