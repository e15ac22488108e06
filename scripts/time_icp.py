"""Times PointFusion / ICPSLAM with ICP odometry (B=8, 640x480) per frame on the GPU box."""
import faulthandler
import os
import sys
import time

faulthandler.dump_traceback_later(int(os.environ.get("GSX_WATCHDOG", "120")), exit=True)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gradslam_b200 as gs
from gradslam_b200.synthetic import make_sequence

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
rgb, depth, K, poses = make_sequence(B, L, 480, 640, seed=0)
frames = gs.RGBDImages(rgb.to(dev), depth.to(dev), K.to(dev), poses.to(dev))
for odom, iters in (("gradicp", 20), ("icp", 20), ("gradicp", 10)):
    slam = gs.PointFusion(odom=odom, numiters=iters, device=dev)
    print("warm-up", odom, iters, flush=True)
    slam(frames)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pc, rec = slam(frames)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = (rec.cpu() - poses).abs().max().item()
    print("PointFusion(odom=%s, numiters=%d) B=%d L=%d: %.1f ms total, %.2f ms per localised frame-step, %.0f frames/s, "
          "max |pose - gt| = %.4f" % (odom, iters, B, L, dt * 1e3, dt * 1e3 / (L - 1), B * L / dt, err), flush=True)
