"""In-process comparison of kernel-tuning variants: generates the bench workload once, then for every
libgsx*.so given on the command line (default: the main build + _lib/variants/*) times whole steps and the
per-kernel breakdown.  Run on the GPU box:  python scripts/tune.py [--L 32] [libs...]"""
import ctypes
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import gradslam_b200 as gs
from gradslam_b200 import _C, profiling
from gradslam_b200.slam.fusionutils import _Workspace
from gradslam_b200.synthetic import make_sequence

args = [a for a in sys.argv[1:]]
opts = {"--L": 32, "--B": 8, "--H": 480, "--W": 640}
for k in list(opts):
    if k in args:
        i = args.index(k)
        opts[k] = int(args[i + 1])
        del args[i:i + 2]
L, B, H, W = opts["--L"], opts["--B"], opts["--H"], opts["--W"]
libs = args or [_C.LIB_PATH] + sorted(glob.glob(os.path.join(ROOT, "gradslam_b200", "_lib", "variants", "*.so")))
dev = torch.device("cuda:0")
t0 = time.time()
rgb, depth, K, poses = make_sequence(B, L, H, W, seed=0)
rgb, depth, K, poses = (t.to(dev) for t in (rgb, depth, K, poses))
print("inputs ready in %.1f s" % (time.time() - t0), flush=True)
frames = gs.RGBDImages(rgb, depth, K, poses)
slam = gs.PointFusion(odom="gt", device=dev)
ref_counts = None
for path in libs:
    handle = ctypes.CDLL(path)
    for name, (res, at) in _C.SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, at
    _C._lib = handle
    _Workspace._cache.clear()
    for _ in range(2):
        pc, _p = slam(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pc, _p = slam(frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    counts = pc.num_points_per_pointcloud.tolist()
    if ref_counts is None:
        ref_counts = counts
    prof, _ = profiling.profile_pointfusion_gt(depth, rgb, K, poses, slam.dist_th, slam.dot_th, slam.sigma)
    prof, finfo = profiling.profile_pointfusion_gt(depth, rgb, K, poses, slam.dist_th, slam.dot_th, slam.sigma)
    line = "%-28s %.3f ms/step  %.0f frames/s  same_counts=%s |" % (os.path.basename(path), ms, B * L / ms * 1e3,
                                                                   counts == ref_counts)
    for name, rows in prof.items():
        tot = sum(r[0] for r in rows)
        line += " %s %.1f us %.0f GB/s |" % (name[:2], 1e3 * tot / len(rows), sum(r[1] for r in rows) / tot / 1e6)
    print(line, flush=True)
print("last frame:", finfo[-1])
