"""Timeline of the pipelined N-GPU bench loop on rank 0 (torch.profiler -> compact event list in gpurun_out/).
torchrun --nproc-per-node 2 scripts/trace_n2.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gradslam_b200 as gs
from gradslam_b200 import parallel
from gradslam_b200.synthetic import make_sequence

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, L, H, W = 8, 32, 480, 640
rgb, depth, K, poses = make_sequence(B, L, H, W, seed=rank)
frames = gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev))
slam = gs.PointFusion(odom="gt", device=dev)
exchange = os.environ.get("TRACE_EXCHANGE", "1") == "1"


def run(steps):
    pending = None
    for _ in range(steps):
        pc, p = slam(frames)
        if pending is not None:
            parallel.gather_maps_end(pending, wait=False)
        pending = parallel.gather_maps_begin(pc) if exchange else None
    if pending is not None:
        parallel.gather_maps_end(pending, wait=True)
    torch.cuda.synchronize(dev)


run(8)
dist.barrier()
torch.cuda.synchronize(dev)
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(6)
if rank == 0:
    path = "gpurun_out/trace_tmp.json"
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    keep = []
    for e in ev:
        if e.get("ph") != "X":
            continue
        cat = e.get("cat", "")
        if cat in ("kernel", "gpu_memcpy", "gpu_memset"):
            keep.append(("G", e["ts"], e["dur"], e.get("args", {}).get("stream"), e["name"][:60]))
        elif cat in ("cuda_runtime", "cuda_driver") and e["dur"] >= 20:
            keep.append(("R", e["ts"], e["dur"], e.get("tid"), e["name"][:60]))
        elif cat == "user_annotation" or (cat == "cpu_op" and e["dur"] >= 200):
            keep.append(("C", e["ts"], e["dur"], e.get("tid"), e["name"][:60]))
    keep.sort(key=lambda r: r[1])
    t0 = keep[0][1]
    with open("gpurun_out/trace_n2_%s.txt" % ("exchange" if exchange else "none"), "w") as f:
        for k, ts, dur, s, name in keep:
            f.write("%s %10.1f %8.1f %s %s\n" % (k, ts - t0, dur, s, name))
    os.remove(path)
dist.destroy_process_group()
