"""BASELINE.json config 3: ICPSLAM 640x480, 10 GN iterations, batch 8, forward + backward (pose-gradient check).
Differentiable mode: K1 forward/backward kernels, CUDA association (frustum tables, exact 1-NN), taped gradLM algebra."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gradslam_b200 as gs
from gradslam_b200.synthetic import make_sequence

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
rgb, depth, K, poses = make_sequence(B, 2, 480, 640, seed=0, yaw0=0.6)  # (the timings below include the upload; see config3_breakdown.py)
d = depth.to(dev).requires_grad_(True)
p = poses.to(dev).requires_grad_(True)
slam = gs.ICPSLAM(odom="gradicp", numiters=10, dsratio=4, device=dev)
for it in range(2):
    d.grad = p.grad = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pc, rec = slam(gs.RGBDImages(rgb.to(dev), d, K.to(dev), p))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    rec.sum().backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("run %d: forward %.1f ms, backward %.1f ms; |d pose/d depth| max %.3e (frame0) %.3e (frame1); "
          "|d/d first pose| max %.3e; pose err vs gt %.2e; finite=%s" % (
              it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, d.grad[:, 0].abs().max().item(), d.grad[:, 1].abs().max().item(),
              p.grad[:, 0].abs().max().item(), (rec.detach().cpu() - poses).abs().max().item(),
              bool(torch.isfinite(d.grad).all() and torch.isfinite(p.grad).all())), flush=True)
with torch.no_grad():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pc2, rec2 = slam(gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev)))
    torch.cuda.synchronize()
    print("fused forward (no grad): %.1f ms; max |pose(fused) - pose(differentiable)| = %.2e" % (
        (time.perf_counter() - t0) * 1e3, (rec2 - rec.detach()).abs().max().item()))
