"""Round-2 layout study: PointFusion(odom='gt') whole-sequence step on the current SoA map store versus the "geo32"
layout (geometry rows (px,py,pz,nx,ny,nz,cc,0) of one 32-byte sector + a separate colour array).
Checks that both layouts give bit-identical maps, then times them.     python scripts/geo32_experiment.py [--L 32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gradslam_b200 as gs
from gradslam_b200 import _C
from gradslam_b200.slam.fusionutils import _Workspace
from gradslam_b200.synthetic import make_sequence

args = sys.argv[1:]
opts = {"--L": 32, "--B": 8, "--H": 480, "--W": 640}
for k in list(opts):
    if k in args:
        opts[k] = int(args[args.index(k) + 1])
L, B, H, W = opts["--L"], opts["--B"], opts["--H"], opts["--W"]
P = H * W
dev = torch.device("cuda:0")
rgb, depth, K, poses = (t.to(dev).contiguous() for t in make_sequence(B, L, H, W, seed=0))
frames = gs.RGBDImages(rgb, depth, K, poses)
slam = gs.PointFusion(odom="gt", device=dev)
lib = _C.lib()
ws = _Workspace.get(dev, B, H, W)
cap = L * P
geo = torch.empty((B, cap, 8), dtype=torch.float32, device=dev)
col = torch.empty((B, cap, 3), dtype=torch.float32, device=dev)
overflow = torch.zeros(1, dtype=torch.int32, device=dev)


def run_geo():
    counts = torch.zeros((2, B), dtype=torch.int32, device=dev)
    rc = lib.gsx_pointfusion_sequence_gt_geo32(
        _C.ptr(geo), _C.ptr(col), _C.ptr(counts), cap, 0, _C.ptr(depth), _C.ptr(rgb), _C.ptr(K), _C.ptr(poses), B, L,
        0, L, H, W, float(slam.dist_th), float(slam.dot_th), float(slam.sigma), _C.ptr(ws.buf), ws.next_epochs(L),
        _C.ptr(overflow), _C.stream_ptr(dev))
    _C.check(rc, "gsx_pointfusion_sequence_gt_geo32")
    return counts[L & 1]


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.cuda.device(dev):
    pc, _ = slam(frames)
    n_geo = run_geo()
    torch.cuda.synchronize()
    n_ref = pc.num_points_per_pointcloud.to(torch.int32)
    assert int(overflow.item()) == 0
    assert torch.equal(n_ref, n_geo), (n_ref.tolist(), n_geo.tolist())
    for b, n in enumerate(n_ref.tolist()):
        assert torch.equal(geo[b, :n, 0:3], pc.points_list[b]), "points differ"
        assert torch.equal(geo[b, :n, 3:6], pc.normals_list[b]), "normals differ"
        assert torch.equal(geo[b, :n, 6:7], pc.features_list[b]), "confidence counts differ"
        assert torch.equal(col[b, :n], pc.colors_list[b]), "colours differ"
    print("geo32 layout: maps bit-identical to the SoA layout, sizes", n_ref.tolist(), flush=True)
    for groups in (1, 2):
        os.environ["GSX_SEQ_GROUPS"] = str(groups)
        ms_soa = timeit(lambda: slam(frames))
        ms_geo = timeit(run_geo)
        print("groups=%d   SoA %.3f ms/step (%.0f frames/s)   geo32 %.3f ms/step (%.0f frames/s)" % (
            groups, ms_soa, B * L / ms_soa * 1e3, ms_geo, B * L / ms_geo * 1e3), flush=True)
