"""GPU twins of the reference's own hot-path tests: the same scenarios, inputs and assertions as
/root/reference/tests/slam/test_fusionutils.py (:305-333 active, :439-483 similar, :879-913 correspondences, :1138-1176
update_map_fusion), tests/odometry/test_icp.py:14-52, test_gradicp.py:14-60, test_icputils.py:284-387 (point_to_plane_ICP
transform recovery, CUDA-only in the reference), :800-867 and :942-1013 (downsampling known answers), restated against this
package with device='cuda:0'.  Inputs are the reference's test data (tests/data/msrd_b2s3, re-packed in
tests/golden/msrd_b2s3.npz by tests/golden/make_golden.py; what tests/common.py::load_test_data returns)."""
import math
import os

import numpy as np
import pytest
import torch
from torch.testing import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_test_data(batch_size=2):
    d = np.load(os.path.join(GOLD, "msrd_b2s3.npz"))
    return tuple(torch.from_numpy(d[k])[:batch_size] for k in ("colors", "depths", "intrinsics", "poses"))


def _frames(batch_size=2):
    import gradslam_b200 as gs

    colors, depths, intrinsics, poses = load_test_data(batch_size)
    return gs.RGBDImages(colors.to(DEV), depths.to(DEV), intrinsics.to(DEV), poses.to(DEV), channels_first=False), colors


def rgbdimages_to_pointclouds(rgbdimages, sigma):
    """the helper of the reference's test file (test_fusionutils.py:15-24)"""
    from gradslam_b200.slam import fusionutils
    from gradslam_b200.structures.utils import pointclouds_from_rgbdimages

    pc_global = pointclouds_from_rgbdimages(rgbdimages)
    pc_local = pointclouds_from_rgbdimages(rgbdimages, global_coordinates=False)
    features = fusionutils.get_alpha(pc_local.points_padded, sigma)
    pc_global.features_padded = (features * pc_global.nonpad_mask.to(features.dtype)).unsqueeze(-1)
    return pc_global


def test_find_active_map_points_twin():
    from gradslam_b200.slam import fusionutils

    rgbd, colors = _frames()
    pc = rgbdimages_to_pointclouds(rgbd[:, 0], 0.6)
    t = fusionutils.find_active_map_points(pc, rgbd[:, 0])
    assert t.shape[0] == rgbd.valid_depth_mask[:, 0].sum()
    projected = torch.zeros_like(colors.to(DEV))
    projected[t[:, 0], 0, t[:, 2], t[:, 3]] = pc.colors_padded[t[:, 0], t[:, 1]]
    assert_close(projected[:, 0:1], colors.to(DEV)[:, 0:1] * rgbd.valid_depth_mask[:, 0:1].float())


def test_find_similar_map_points_twin():
    from gradslam_b200.slam import fusionutils

    rgbd, _ = _frames()
    dist_th, dot_th = 0.05 ** 0.5, 0.9
    pc = rgbdimages_to_pointclouds(rgbd[:, 0], 0.6)
    active = fusionutils.find_active_map_points(pc, rgbd[:, 0])
    similar, is_similar = fusionutils.find_similar_map_points(pc, rgbd[:, 0], active, dist_th, dot_th)
    # only points with zero normals (despite valid depths) are removed
    not_similar = active[is_similar == False]  # noqa: E712
    frame_normals = torch.zeros_like(pc.normals_padded)
    frame_normals[not_similar[:, 0], not_similar[:, 1]] = rgbd.normal_map[
        not_similar[:, 0], 0, not_similar[:, 2], not_similar[:, 3]]
    assert frame_normals.abs().max() == 0
    assert active.shape[0] - similar.shape[0] == (
        pc.normals_list[0].eq(0).all(-1).sum() + pc.normals_list[1].eq(0).all(-1).sum()).item()
    assert pc.points_list[0].eq(0).all(-1).sum().item() == 0
    assert pc.points_list[1].eq(0).all(-1).sum().item() == 0


def test_find_correspondences_twin():
    from gradslam_b200.slam import fusionutils

    rgbd, _ = _frames()
    pc = rgbdimages_to_pointclouds(rgbd[:, 0], 0.6)
    t = fusionutils.find_correspondences(pc, rgbd[:, 0], 0.05 ** 0.5, 0.9)
    num_valid = rgbd.valid_depth_mask[:, 0].sum()
    valid_zero_normals = (rgbd.normal_map[:, 0].eq(0).all(-1).int()
                          - (rgbd.valid_depth_mask[:, 0] == False).squeeze(-1).int())  # noqa: E712
    assert valid_zero_normals.abs().sum() == valid_zero_normals.sum()
    assert (rgbd.vertex_map[:, 0].eq(0).all(-1).int()
            - (rgbd.valid_depth_mask[:, 0] == False).squeeze(-1).int()).abs().sum() == 0  # noqa: E712
    assert t.shape[0] == num_valid - valid_zero_normals.sum()


def test_update_map_fusion_twin():
    from gradslam_b200.slam import fusionutils

    rgbd, _ = _frames()
    pc = rgbdimages_to_pointclouds(rgbd[:, 0], 0.6)
    n0 = pc.num_points_per_pointcloud
    pc = fusionutils.update_map_fusion(pc, rgbd[:, 1], 0.05 ** 0.5, 0.9, 0.6)
    n1 = pc.num_points_per_pointcloud
    assert n1.gt(n0).all()
    # parameters under which more points fuse
    pc2 = rgbdimages_to_pointclouds(rgbd[:, 0], 0.6)
    m0 = pc2.num_points_per_pointcloud
    pc2 = fusionutils.update_map_fusion(pc2, rgbd[:, 1], 0.4 ** 0.5, 0.5, 0.6)
    m1 = pc2.num_points_per_pointcloud
    assert m1.gt(m0).all()
    assert n1.gt(m1).all()


def _rigid_case(axis="z", rad=0.1):
    from gradslam_b200.structures.utils import pointclouds_from_rgbdimages

    rgbd, colors = _frames(1)
    src = pointclouds_from_rgbdimages(rgbd[:, 0])
    c, s = math.cos(rad), math.sin(rad)
    if axis == "z":
        T = [[c, -s, 0.0, 0.05], [s, c, 0.0, 0.03], [0.0, 0.0, 1.0, 0.01], [0.0, 0.0, 0.0, 1.0]]
    else:
        T = [[1.0, 0.0, 0.0, 0.05], [0.0, c, -s, 0.03], [0.0, s, c, 0.01], [0.0, 0.0, 0.0, 1.0]]
    T = torch.tensor(T, device=DEV, dtype=colors.dtype)
    return src, src.transform(T), T


@pytest.mark.parametrize("which", ["icp", "gradicp"])
def test_odometry_provider_recovers_transform_twin(which):
    """test_icp.py:14-52 / test_gradicp.py:14-60: 30 iterations, dist_thresh 0.2, default assert_allclose tolerances"""
    from gradslam_b200.odometry.gradicp import GradICPOdometryProvider
    from gradslam_b200.odometry.icp import ICPOdometryProvider

    src, tgt, T = _rigid_case("z", 0.1)
    if which == "icp":
        odom = ICPOdometryProvider(numiters=30, damp=1e-8, dist_thresh=0.2)
    else:
        odom = GradICPOdometryProvider(numiters=30, damp=1e-8, dist_thresh=0.2, lambda_max=2.0, B=1.0, B2=1.0, nu=200.0)
    out = odom.provide(tgt, src).squeeze(1).squeeze(0)
    assert out.shape == T.shape
    assert_close(out, T, rtol=1e-4, atol=1e-5)  # (torch.testing.assert_allclose defaults for float32)


def test_point_to_plane_icp_recovers_transform_twin():
    """test_icputils.py:284-387 (CUDA-only in the reference): 100 iterations, no distance threshold"""
    from gradslam_b200.odometry.icputils import point_to_plane_ICP

    src, tgt, T = _rigid_case("x", 0.2)
    t, idx = point_to_plane_ICP(src.points_padded, tgt.points_padded, tgt.normals_padded, torch.eye(4, device=DEV), 100,
                                1e-8, None)
    assert t.shape == T.shape
    assert_close(t, T, rtol=1e-4, atol=1e-5)


def test_downsample_pointclouds_twin():
    import gradslam_b200 as gs
    from gradslam_b200.odometry.icputils import downsample_pointclouds

    points = torch.tensor([[5.0, 5.0, 5.0], [3.0, 3.0, 3.0], [1.0, 2.0, 3.0], [3.0, 2.0, 1.0], [1.0, 0.0, 1.0],
                           [0.0, 0.0, 0.0]], device=DEV).unsqueeze(0)
    normals, colors = points * -1, points * 2
    table = torch.tensor([[0, 0, 0, 0], [0, 1, 4, 2], [0, 2, 3, 1], [0, 3, 0, 3], [0, 4, 3, 3], [0, 5, 3, 6]],
                         device=DEV, dtype=torch.int64)
    ds = downsample_pointclouds(gs.Pointclouds(points, normals, colors), table, 3)
    want = torch.tensor([[5.0, 5.0, 5.0], [3.0, 2.0, 1.0], [1.0, 0.0, 1.0], [0.0, 0.0, 0.0]], device=DEV).unsqueeze(0)
    assert ds.points_padded.shape == want.shape
    assert_close(ds.points_padded, want)
    assert_close(ds.normals_padded, want * -1)
    assert_close(ds.colors_padded, want * 2)
    ds = downsample_pointclouds(gs.Pointclouds(points), table, 2)
    want = torch.tensor([[5.0, 5.0, 5.0], [3.0, 3.0, 3.0]], device=DEV).unsqueeze(0)
    assert ds.points_padded.shape == want.shape and ds.normals_padded is None
    assert_close(ds.points_padded, want)


def test_downsample_rgbdimages_twin():
    import gradslam_b200 as gs
    from gradslam_b200.odometry.icputils import downsample_rgbdimages

    image = torch.arange(12, dtype=torch.float32, device=DEV).view(1, 1, 3, 4, 1).repeat(1, 1, 1, 1, 3)
    depth = torch.ones_like(image[..., :1])
    eye = torch.eye(4, device=DEV).unsqueeze(0).unsqueeze(0)
    rgbd = gs.RGBDImages(image, depth, eye, eye, channels_first=False)
    ds = downsample_rgbdimages(rgbd, 2)
    want_p = torch.tensor([[0.0, 0.0, 1.0], [2.0, 0.0, 1.0], [0.0, 2.0, 1.0], [2.0, 2.0, 1.0]], device=DEV).unsqueeze(0)
    want_c = torch.tensor([[0.0] * 3, [2.0] * 3, [8.0] * 3, [10.0] * 3], device=DEV).unsqueeze(0)
    want_n = rgbd.normal_map[..., ::2, ::2, :].reshape(1, ds.normals_padded.shape[1], 3)
    assert ds.points_padded.shape == want_p.shape
    assert_close(ds.points_padded, want_p, rtol=1e-4, atol=1e-5)  # (closed-form K^-1 adds 1e-6 to fx, fy)
    assert_close(ds.colors_padded, want_c)
    assert_close(ds.normals_padded, want_n)
