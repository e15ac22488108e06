"""Small PointFusion(odom='gt') run for profilers: B=8, 640x480, L frames (default 10), one slam() call.
Usage (on the GPU box, one GPU):
  ncu --set full --clock-control none --import-source on -k regex:k_merge_append -s 8 -c 1 -o gpurun_out/k4 \
      python scripts/ncu_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gradslam_b200 as gs
from gradslam_b200.synthetic import make_sequence

L = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
H = int(sys.argv[3]) if len(sys.argv) > 3 else 480
W = int(sys.argv[4]) if len(sys.argv) > 4 else 640
dev = torch.device("cuda:0")
rgb, depth, K, poses = make_sequence(B, L, H, W, seed=0)
frames = gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev))
slam = gs.PointFusion(odom="gt", device=dev)
pc, _ = slam(frames)
torch.cuda.synchronize()
print("points per map:", pc.num_points_per_pointcloud.tolist())
