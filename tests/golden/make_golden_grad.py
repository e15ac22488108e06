"""Freezes GRADIENTS of the UNMODIFIED gradslam reference (PyTorch autograd through its ATen op chains) as fixtures
(run in the build container only):

    python tests/golden/make_golden_grad.py     ->  tests/golden/ref_grad.npz

The reference's own tests pin no backward pass except get_alpha; these vectors pin the oracle's
autograd (tests/test_oracle_golden.py), which in turn is what the CUDA backward kernels are compared with on the GPU
(tests/test_gpu_backward.py).  Inputs are NOT stored: the tests regenerate them from the recorded seeds.

Cases (all float32, CPU):
  pf_gt       PointFusion(odom='gt'), B=1, L=2, 24x32: d(sum w.map points/colours/ccounts) / d(depth, rgb)
  gradicp     point_to_plane_gradICP, 4 iterations on a 40x56 frame cloud: d(sum w.T) / d(source cloud)
  icp         point_to_plane_ICP, same
  icpslam     ICPSLAM(odom='gradicp', numiters=3, dsratio=2), B=1, L=2, 32x40: d(sum w.poses) / d(depth)
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from ref_loader import load_reference  # noqa: E402

load_reference()
from gradslam.geometry.se3utils import se3_exp  # noqa: E402
from gradslam.odometry.icputils import point_to_plane_gradICP, point_to_plane_ICP  # noqa: E402
from gradslam.slam.icpslam import ICPSLAM  # noqa: E402
from gradslam.slam.pointfusion import PointFusion  # noqa: E402
from gradslam.structures.rgbdimages import RGBDImages  # noqa: E402

from gradslam_b200.synthetic import make_sequence  # noqa: E402

GRAD_CASES = {
    "pf_gt": dict(B=1, L=2, H=24, W=32, seed=41, wseed=5),
    "icp_cloud": dict(H=40, W=56, seed=31, wseed=1, numiters=4, xi=[0.01, -0.005, 0.008, 0.01, -0.01, 0.005]),
    "icpslam": dict(B=1, L=2, H=32, W=40, seed=17, wseed=9, numiters=3, dsratio=2),
}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)  # reduction order of the reference's ATen sums must not depend on the pool
    out = {}

    c = GRAD_CASES["pf_gt"]
    rgb, depth, K, poses = make_sequence(c["B"], c["L"], c["H"], c["W"], seed=c["seed"], yaw0=0.6)
    d, col = depth.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    pc, _ = PointFusion(odom="gt")(RGBDImages(col, d, K, poses))
    n = int(pc.num_points_per_pointcloud[0])
    g = torch.Generator().manual_seed(c["wseed"])
    wp, wc, wf = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g)
    ((pc.points_list[0] * wp).sum() + (pc.colors_list[0] * wc).sum() + (pc.features_list[0] * wf).sum()).backward()
    out["pf_gt/count"] = np.array([n])
    out["pf_gt/d_depth"] = d.grad.numpy()
    out["pf_gt/d_rgb"] = col.grad.numpy()
    print("pf_gt: n=%d |d_depth| max %.3e |d_rgb| max %.3e" % (n, d.grad.abs().max(), col.grad.abs().max()))

    c = GRAD_CASES["icp_cloud"]
    rgb, depth, K, poses = make_sequence(1, 1, c["H"], c["W"], seed=c["seed"], hole_fraction=0.0, yaw0=0.6)
    f = RGBDImages(rgb, depth, K, poses)
    tgt = f.global_vertex_map[0, 0].reshape(-1, 3).contiguous()
    tgt_n = f.global_normal_map[0, 0].reshape(-1, 3).contiguous()
    T_true = se3_exp(torch.tensor(c["xi"]).view(6, 1))
    src0 = (tgt @ T_true[:3, :3].t() + T_true[:3, 3]).contiguous()
    w = torch.randn(4, 4, generator=torch.Generator().manual_seed(c["wseed"]))
    for name, fn in (("gradicp", point_to_plane_gradICP), ("icp", point_to_plane_ICP)):
        s = src0.clone().requires_grad_(True)
        T, _ = fn(s.unsqueeze(0), tgt.unsqueeze(0), tgt_n.unsqueeze(0), torch.eye(4), numiters=c["numiters"])
        (T * w).sum().backward()
        out[name + "/T"] = T.detach().numpy()
        out[name + "/d_src"] = s.grad.numpy()
        print("%s: |d_src| max %.3e" % (name, s.grad.abs().max()))

    c = GRAD_CASES["icpslam"]
    rgb, depth, K, poses = make_sequence(c["B"], c["L"], c["H"], c["W"], seed=c["seed"], yaw0=0.6)
    d = depth.clone().requires_grad_(True)
    _, rec = ICPSLAM(odom="gradicp", numiters=c["numiters"], dsratio=c["dsratio"])(RGBDImages(rgb, d, K, poses))
    w = torch.randn(rec.shape, generator=torch.Generator().manual_seed(c["wseed"]))
    (rec * w).sum().backward()
    out["icpslam/poses"] = rec.detach().numpy()
    out["icpslam/d_depth"] = d.grad.numpy()
    print("icpslam: |d_depth| max %.3e, finite %s" % (d.grad.abs().max(), bool(torch.isfinite(d.grad).all())))

    np.savez_compressed(os.path.join(HERE, "ref_grad.npz"), **out)
    print("wrote ref_grad.npz (%d arrays)" % len(out))


if __name__ == "__main__":
    main()
