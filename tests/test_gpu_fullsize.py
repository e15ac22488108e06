"""Full-size (640x480, BASELINE.json's metric size) checks of the CUDA path: a short oracle comparison plus
size-independent properties (append-all on an empty map, re-fusing the same frame, determinism, step == sequence)."""
import math

import pytest
import torch

import gsx_oracle as oracle
from gradslam_b200.synthetic import make_sequence

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, W = 480, 640
DOT_TH = math.cos(20 * math.pi / 180)


def _frames(gs, rgb, depth, K, poses):
    return gs.RGBDImages(rgb.to(DEV), depth.to(DEV), K.to(DEV), poses.to(DEV))


def test_fullsize_sequence_matches_oracle():
    import gradslam_b200 as gs

    rgb, depth, K, poses = make_sequence(1, 3, H, W, seed=0)
    pc, _ = gs.PointFusion(odom="gt", device=DEV)(_frames(gs, rgb, depth, K, poses))
    ref = oracle.run_slam(rgb, depth, K, poses, odom="gt")
    assert pc.num_points_per_pointcloud.tolist() == ref.map.counts()
    # canonical arithmetic end to end (the confidence weight's exp is taken in double on both sides): the fused map is
    # BIT-identical to the oracle's, 363 k surfels after three frames
    assert torch.equal(pc.points_list[0].cpu(), ref.map.points[0])
    assert torch.equal(pc.normals_list[0].cpu(), ref.map.normals[0])
    assert torch.equal(pc.colors_list[0].cpu(), ref.map.colors[0])
    assert torch.equal(pc.features_list[0].cpu(), ref.map.ccounts[0])


def test_fullsize_run_matches_frozen_reference():
    """BASELINE.json's frame size and the bench's input distribution (2 % random holes) against the UNMODIFIED
    reference, frozen by tests/golden/make_golden.py (640x480, B=1, L=6, odom='gt'): map size after every frame within
    1e-4 (measured and asserted: -1 / -3 / -2 points of ~4e5 after frames 4-6), checksums, and a 1-in-53 sample of the
    final surfels within north_star's 1e-3 (tests/golden/fullsize.py states every bound)."""
    import os

    import numpy as np

    import gradslam_b200 as gs
    from golden.fullsize import FULL_L, check_against_frozen_reference

    ref = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_slam.npz")))
    rgb, depth, K, poses = make_sequence(1, FULL_L, H, W, seed=0)
    frames = _frames(gs, rgb, depth, K, poses)
    slam = gs.PointFusion(odom="gt", device=DEV)
    pc = gs.Pointclouds(device=DEV)
    sizes = []
    for s in range(FULL_L):
        pc, _ = slam.step(pc, frames[:, s], None, inplace=True)
        sizes.append(int(pc.num_points_per_pointcloud[0]))
    check_against_frozen_reference(ref, sizes, pc.points_list[0], pc.normals_list[0], pc.colors_list[0],
                                   pc.features_list[0])
    # and the whole-sequence call gives the same map
    pc2, _ = slam(frames)
    assert int(pc2.num_points_per_pointcloud[0]) == sizes[-1]
    assert torch.equal(pc2.points_list[0], pc.points_list[0])


def test_frame_maps_equal_frozen_reference_run():
    """K1 on random-hole input: local vertex / normal maps BIT-identical to the reference's CPU run (FMA cross product
    and norm, see gsx_common.cuh cross_ref / norm_ref), global maps within an ulp (reference: BLAS-ordered einsum)."""
    import os

    import numpy as np

    import gradslam_b200 as gs

    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_slam.npz"))
    rgb, depth, K, poses = make_sequence(2, 2, 60, 80, seed=6)
    fr = _frames(gs, rgb, depth, K, poses)
    assert torch.equal(fr.vertex_map.cpu(), torch.from_numpy(ref["k1/vertex"]))
    assert torch.equal(fr.normal_map.cpu(), torch.from_numpy(ref["k1/normal"]))
    torch.testing.assert_close(fr.global_vertex_map.cpu(), torch.from_numpy(ref["k1/gvertex"]), rtol=0, atol=1e-6)
    torch.testing.assert_close(fr.global_normal_map.cpu(), torch.from_numpy(ref["k1/gnormal"]), rtol=0, atol=2.5e-7)


def test_fullsize_properties():
    import gradslam_b200 as gs
    from gradslam_b200.slam import fusionutils as fu

    B = 4
    rgb, depth, K, poses = make_sequence(B, 4, H, W, seed=9)
    frames = _frames(gs, rgb, depth, K, poses)
    valid = depth[:, :, :, :, 0] > 0
    # (1) fusing a frame into an empty map appends exactly its valid pixels, in row-major order
    pc = fu.update_map_fusion(gs.Pointclouds(device=DEV), frames[:, 0], 0.05, DOT_TH, 0.6)
    assert pc.num_points_per_pointcloud.tolist() == valid[:, 0].flatten(1).sum(1).tolist()
    gv = frames[:, 0].global_vertex_map[:, 0]
    for b in range(B):
        assert torch.equal(pc.points_list[b], gv[b][valid[b, 0].to(DEV)])
    # (2) fusing the SAME frame again: every valid pixel with a non-zero normal merges with its own surfel (position
    #     unchanged up to rounding, confidence doubles), only zero-normal pixels are appended again
    gn = frames[:, 0].global_normal_map[:, 0]
    zero_n = ((gn.abs().sum(-1) == 0).cpu() & valid[:, 0]).flatten(1).sum(1)
    pc2 = fu.update_map_fusion(pc, frames[:, 0], 0.05, DOT_TH, 0.6)
    n1 = pc.num_points_per_pointcloud.cpu()
    assert (pc2.num_points_per_pointcloud.cpu() == n1 + zero_n).all()
    for b in range(B):
        k = int(n1[b])
        torch.testing.assert_close(pc2.points_list[b][:k], pc.points_list[b], rtol=1e-6, atol=1e-6)
        merged = gn[b][valid[b, 0].to(DEV)].abs().sum(-1) > 0
        torch.testing.assert_close(pc2.features_list[b][:k][merged], 2 * pc.features_list[b][merged], rtol=1e-6, atol=0)
        assert torch.equal(pc2.features_list[b][:k][~merged], pc.features_list[b][~merged])
    # (3) determinism (the atomic arg-min is order independent) and sequence call == step calls, bit for bit
    slam = gs.PointFusion(odom="gt", device=DEV)
    a, _ = slam(frames)
    b_, _ = slam(frames)
    c = gs.Pointclouds(device=DEV)
    for s in range(4):
        c, _ = slam.step(c, frames[:, s], None, inplace=True)
    assert a.num_points_per_pointcloud.tolist() == b_.num_points_per_pointcloud.tolist() == c.num_points_per_pointcloud.tolist()
    for i in range(B):
        for attr in ("points_list", "normals_list", "colors_list", "features_list"):
            assert torch.equal(getattr(a, attr)[i], getattr(b_, attr)[i])
            assert torch.equal(getattr(a, attr)[i], getattr(c, attr)[i])
    # (4) monotone bookkeeping: sizes never shrink, confidence counts are positive, padding rows are zero
    assert (a.num_points_per_pointcloud.cpu() >= n1).all()
    assert (a.features_padded[a.nonpad_mask] > 0).all()
    assert a.points_padded[~a.nonpad_mask].abs().sum() == 0


def test_fullsize_icp_localisation_matches_oracle():
    """One ICP-localised step at 640x480 (dsratio 4 => 19 200 source points, grid 1-NN): pose within 1e-4."""
    import gradslam_b200 as gs

    rgb, depth, K, poses = make_sequence(1, 2, H, W, seed=3, yaw0=0.6)
    slam = gs.PointFusion(odom="gradicp", numiters=5, device=DEV)
    pc, rec = slam(_frames(gs, rgb, depth, K, poses))
    ref = oracle.run_slam(rgb, depth, K, poses, odom="gradicp", numiters=5)
    torch.testing.assert_close(rec.cpu(), ref.poses, rtol=0, atol=1e-4)
    assert abs(pc.num_points_per_pointcloud.tolist()[0] - ref.map.counts()[0]) <= ref.map.counts()[0] // 1000
