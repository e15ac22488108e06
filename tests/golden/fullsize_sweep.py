"""How far the oracle's full-size map is from the UNMODIFIED reference's on seeds other than the frozen one (build
container only; nothing is stored as a fixture - the output is pasted into fullsize_sweep.txt).

    python tests/golden/fullsize_sweep.py [seed ...]

Per seed: PointFusion(odom='gt'), 640x480, B=1, L=6, default 2 % random holes; map size after every frame on both sides,
and for the final maps the nearest-neighbour distance from every reference surfel to ours."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.simplefilter("ignore")

from ref_loader import load_reference  # noqa: E402

load_reference()
from gradslam.slam.pointfusion import PointFusion  # noqa: E402
from gradslam.structures.pointclouds import Pointclouds  # noqa: E402
from gradslam.structures.rgbdimages import RGBDImages  # noqa: E402
from scipy.spatial import cKDTree  # noqa: E402

import gsx_oracle as oracle  # noqa: E402
from gradslam_b200.synthetic import make_sequence  # noqa: E402

L = 6
seeds = [int(s) for s in sys.argv[1:]] or [1, 2, 3]
for seed in seeds:
    rgb, depth, K, poses = make_sequence(1, L, 480, 640, seed=seed)
    frames = RGBDImages(rgb, depth, K, poses)
    slam = PointFusion(odom="gt")
    pc = Pointclouds()
    ref_sizes = []
    for s in range(L):
        pc, _ = slam.step(pc, frames[:, s], None, inplace=True)
        ref_sizes.append(int(pc.num_points_per_pointcloud[0]))
    ours = oracle.run_slam(rgb, depth, K, poses, odom="gt", record_sizes=True) if "record_sizes" in \
        oracle.run_slam.__code__.co_varnames else None
    if ours is None:  # sizes per frame through the step-wise oracle
        sizes = []
        for n in range(1, L + 1):
            r = oracle.run_slam(rgb[:, :n], depth[:, :n], K, poses[:, :n], odom="gt")
            sizes.append(r.map.counts()[0])
        final = r.map
    else:
        sizes, final = ours.sizes, ours.map
    delta = [a - b for a, b in zip(sizes, ref_sizes)]
    ref_pts = pc.points_list[0].numpy()
    dist, _ = cKDTree(final.points[0].numpy()).query(ref_pts, k=1)
    print("seed %d: reference sizes %s | ours - reference %s | reference surfels farther than 2e-5 from ours: %d of %d "
          "(max %.2e), farther than 1e-3: %d" % (seed, ref_sizes, delta, int((dist > 2e-5).sum()), dist.size,
                                                 dist.max(), int((dist > 1e-3).sum())), flush=True)
