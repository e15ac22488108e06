"""Stage-by-stage timing of one ICP-localised step (diagnostic)."""
import faulthandler
import os
import sys
import time

faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gradslam_b200 as gs
from gradslam_b200.synthetic import make_sequence

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 640)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
rgb, depth, K, poses = make_sequence(B, 3, H, W, seed=0)
frames = gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev))
slam = gs.PointFusion(odom="gradicp", numiters=iters, device=dev)


def tick(msg, t0):
    torch.cuda.synchronize()
    print("%-40s %.1f ms" % (msg, (time.perf_counter() - t0) * 1e3), flush=True)


pc = gs.Pointclouds(device=dev)
t0 = time.perf_counter()
f0 = frames[:, 0]
pc, p0 = slam.step(pc, f0, None, inplace=True)
tick("frame 0 (map only)", t0)
print("counts", pc.num_points_per_pointcloud.tolist(), "bound", pc._bound, "cap", pc.capacity, flush=True)
prev = f0
for s in (1, 2):
    live = frames[:, s]
    t0 = time.perf_counter()
    pose = slam._localize(pc, live, prev)
    tick("frame %d localize" % s, t0)
    live.poses = pose
    t0 = time.perf_counter()
    pc = slam._map(pc, live, True)
    tick("frame %d map" % s, t0)
    print("  pose err", (pose[:, 0].cpu() - poses[:, s]).abs().max().item(), flush=True)
    prev = live
print("done", flush=True)
