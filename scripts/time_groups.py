"""Times PointFusion(odom='gt') whole-sequence steps for every batch-group count of the sequence driver
(GSX_SEQ_GROUPS = 1..4) and checks that the fused maps are bit-identical.  python scripts/time_groups.py [--L 32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gradslam_b200 as gs
from gradslam_b200.synthetic import make_sequence

args = sys.argv[1:]
opts = {"--L": 32, "--B": 8, "--H": 480, "--W": 640}
for k in list(opts):
    if k in args:
        opts[k] = int(args[args.index(k) + 1])
L, B, H, W = opts["--L"], opts["--B"], opts["--H"], opts["--W"]
dev = torch.device("cuda:0")
rgb, depth, K, poses = (t.to(dev) for t in make_sequence(B, L, H, W, seed=0))
frames = gs.RGBDImages(rgb, depth, K, poses)
slam = gs.PointFusion(odom="gt", device=dev)
ref = None
for G in (1, 2, 3, 4, 2, 1):
    os.environ["GSX_SEQ_GROUPS"] = str(G)
    for _ in range(3):
        pc, _p = slam(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pc, _p = slam(frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    cur = [pc.num_points_per_pointcloud.clone()] + [getattr(pc, k + "_padded").clone() for k in
                                                    ("points", "normals", "colors", "features")]
    same = True if ref is None else all(torch.equal(a, b) for a, b in zip(ref, cur))
    if ref is None:
        ref = cur
    print("groups=%d  %.3f ms/step  %.0f frames/s  identical_to_groups1=%s" % (G, ms, B * L / ms * 1e3, same), flush=True)
