"""Kernel micro-benchmark on a FIXED state: fuses frames [0, S) with the main build, then times K1r / K2 / K4 of frame S
for every libgsx*.so given (default: main + _lib/variants/*) on identical inputs - K2 on the same map and frame records,
K4 on a scratch copy of the map with the main build's arg-min records - so timing-only ablations that produce wrong
results do not change the workload.   python scripts/kmicro.py [--S 16] [--B 8] [--reps 20] [libs...]"""
import ctypes
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import gradslam_b200 as gs
from gradslam_b200 import _C
from gradslam_b200.slam.fusionutils import _Workspace
from gradslam_b200.synthetic import make_sequence

args = list(sys.argv[1:])
opts = {"--S": 16, "--B": 8, "--H": 480, "--W": 640, "--reps": 20}
for k in list(opts):
    if k in args:
        i = args.index(k)
        opts[k] = int(args[i + 1])
        del args[i:i + 2]
S, B, H, W, reps = opts["--S"], opts["--B"], opts["--H"], opts["--W"], opts["--reps"]
libs = args or [_C.LIB_PATH] + sorted(glob.glob(os.path.join(ROOT, "gradslam_b200", "_lib", "variants", "*.so")))
dev = torch.device("cuda:0")
P = H * W
rgb, depth, K, poses = make_sequence(B, S + 1, H, W, seed=0)
rgb, depth, K, poses = (t.to(dev) for t in (rgb, depth, K, poses))
slam = gs.PointFusion(odom="gt", device=dev)
pc, _ = slam(gs.RGBDImages(rgb[:, :S].contiguous(), depth[:, :S].contiguous(), K, poses[:, :S].contiguous()))
counts = pc.num_points_per_pointcloud.tolist()
M = sum(counts)
pc.reserve(max(counts) + P)
main = _C.lib()
ws = _Workspace.get(dev, B, H, W)
stream = _C.stream_ptr(dev)
d_s = depth[:, S].contiguous()
c_s = rgb[:, S].contiguous()
p_s = poses[:, S].contiguous()
cin = pc._counts_dev[pc._cur]
cout = torch.zeros_like(cin)
ovf = torch.zeros(1, dtype=torch.int32, device=dev)
geo2, col2 = torch.empty_like(pc._geo), torch.empty_like(pc._col)
L2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def load(path):
    h = ctypes.CDLL(path)
    for name, (res, at) in _C.SIGNATURES.items():
        fn = getattr(h, name)
        fn.restype, fn.argtypes = res, at
    return h


def k1(lib):
    _C.check(lib.gsx_fusion_frame_records(_C.ptr(d_s), P, _C.ptr(K), 16, _C.ptr(p_s), 16, None, None, None, B, H, W, 0.6,
                                          _C.ptr(ws.buf), stream), "k1")


def k2(lib):
    _C.check(lib.gsx_fusion_project_select(_C.ptr(pc._geo), _C.ptr(cin), pc.capacity, max(counts), _C.ptr(p_s), 16,
                                           _C.ptr(K), 16, B, H, W, 0.05, slam.dot_th, _C.ptr(ws.buf), stream), "k2")


def k4(lib):
    _C.check(lib.gsx_fusion_merge_append(_C.ptr(geo2), _C.ptr(col2), 1, _C.ptr(cin), _C.ptr(cout), pc.capacity,
                                         _C.ptr(c_s), P * 3, B, H, W, _C.ptr(ws.buf), _C.ptr(ovf), None, stream), "k4")


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


print("state: B=%d frame %d, map rows %d (%.0f MB of geometry rows), P*B=%d" % (B, S, M, M * 32 / 1e6, P * B), flush=True)
for path in libs:
    lib = load(path)
    t1 = t2 = t4 = 0.0
    for r in range(reps + 2):
        L2.zero_()  # flush L2 between repetitions
        a = timed(lambda: k1(lib))
        b_ = timed(lambda: k2(lib))
        k1(main)  # K4 consumes the MAIN build's records on a scratch copy of the map
        k2(main)
        geo2.copy_(pc._geo)
        col2.copy_(pc._col)
        L2.zero_()
        c = timed(lambda: k4(lib))
        if r >= 2:
            t1, t2, t4 = t1 + a, t2 + b_, t4 + c
    print("%-28s K1r %6.1f us | K2 %6.1f us | K4 %6.1f us | new counts %s" % (
        os.path.basename(path), t1 / reps, t2 / reps, t4 / reps, (cout - cin).tolist()[:2]), flush=True)
