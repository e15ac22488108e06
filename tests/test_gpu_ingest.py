"""Dataset-native ingest: uint8 colour + uint16 depth -> float32 on the device, bit-identical to the host-side
conversion of the reference's loaders (gradslam/datasets/icl.py:467-513)."""
import numpy as np
import pytest
import torch

from gradslam_b200.synthetic import make_sequence

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _raw(B, L, H, W, seed):
    rgb, depth, K, poses = make_sequence(B, L, H, W, seed=seed)
    col = (rgb.numpy() * 255.0).astype(np.uint8)
    dep = np.round(depth.numpy()[..., 0] * 5000.0).astype(np.uint16)
    return col, dep, K, poses


@pytest.mark.parametrize("shape,normalize", [((2, 2, 24, 32), False), ((1, 3, 17, 23), True)])
def test_raw_to_float_matches_loader_arithmetic(shape, normalize):
    from gradslam_b200 import ingest

    col, dep, K, poses = _raw(*shape, seed=1)
    rgb, depth = ingest.raw_to_float(torch.from_numpy(col).to(DEV), torch.from_numpy(dep).to(DEV), 5000.0, normalize)
    want_rgb = col.astype(float)  # np.asarray(imread(...), dtype=float)
    if normalize:
        want_rgb = want_rgb / 255.0
    want_depth = dep.astype(float)[..., None] / 5000.0
    assert torch.equal(rgb.cpu(), torch.from_numpy(want_rgb).float())
    assert torch.equal(depth.cpu(), torch.from_numpy(want_depth).float())


def test_pointfusion_on_raw_input_equals_float_input():
    import gradslam_b200 as gs
    from gradslam_b200 import ingest

    B, L, H, W = 2, 5, 48, 64
    col, dep, K, poses = _raw(B, L, H, W, seed=2)
    raw = ingest.RawRGBD(torch.from_numpy(col).pin_memory(), torch.from_numpy(dep).pin_memory(), K, poses)
    slam = gs.PointFusion(odom="gt", device=DEV)
    pc_raw, p_raw = slam(raw)
    fl = ingest.rgbdimages_from_raw(torch.from_numpy(col), torch.from_numpy(dep), K, poses, device=DEV)
    host_rgb = torch.from_numpy(col.astype(float)).float()
    host_depth = torch.from_numpy(dep.astype(float)[..., None] / 5000.0).float()
    assert torch.equal(fl.rgb_image.cpu(), host_rgb) and torch.equal(fl.depth_image.cpu(), host_depth)
    pc_f, p_f = slam(gs.RGBDImages(host_rgb.to(DEV), host_depth.to(DEV), K.to(DEV), poses.to(DEV)))
    assert pc_raw.num_points_per_pointcloud.tolist() == pc_f.num_points_per_pointcloud.tolist()
    for b in range(B):
        assert torch.equal(pc_raw.points_list[b], pc_f.points_list[b])
        assert torch.equal(pc_raw.colors_list[b], pc_f.colors_list[b])
        assert torch.equal(pc_raw.features_list[b], pc_f.features_list[b])
    assert torch.equal(p_raw, p_f)
    with pytest.raises(ValueError):
        gs.PointFusion(odom="gradicp", device=DEV)(raw)
    with pytest.raises(TypeError):
        ingest.RawRGBD(torch.zeros(1, 1, 4, 4, 3), torch.zeros(1, 1, 4, 4, dtype=torch.int16), K, poses)


def test_calibration_contract_matches_frozen_reference():
    """scale_intrinsics (datasets/datautils.py:73-122) and frame-0-relative poses (datasets/icl.py:515-533) on the device
    against values frozen from the reference (tests/golden/make_golden.py, `f2/*`): the intrinsics are float32 products
    and must be bit-identical; the relative poses go through a general 4x4 inverse (LAPACK LU in the reference,
    Gauss-Jordan here) and are held to 1e-5."""
    import os

    import numpy as np

    from gradslam_b200 import ingest

    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_slam.npz"))
    K = torch.from_numpy(ref["f2/K_in"]).to(DEV)
    hr, wr = (float(x) for x in ref["f2/ratios"])
    assert torch.equal(ingest.scale_intrinsics(K, hr, wr).cpu(), torch.from_numpy(ref["f2/K_scaled"]))
    assert torch.equal(ingest.scale_intrinsics(K[:, :3, :3].contiguous(), 0.5, 0.75).cpu(),
                       torch.from_numpy(ref["f2/K3_scaled"]))
    # the host mirror gives the same bits
    assert torch.equal(ingest.scale_intrinsics(K.cpu(), hr, wr), torch.from_numpy(ref["f2/K_scaled"]))
    poses = torch.from_numpy(ref["f2/poses_abs"]).to(DEV)
    rel = ingest.relative_poses(poses)
    torch.testing.assert_close(rel.cpu(), torch.from_numpy(ref["f2/poses_rel"]), rtol=1e-5, atol=1e-5)
    assert torch.equal(rel[:, 0, 3].cpu(), torch.tensor([[0.0, 0.0, 0.0, 1.0]] * 2))
    torch.testing.assert_close(rel[:, 0].cpu(), torch.eye(4).repeat(2, 1, 1), rtol=0, atol=1e-5)
    torch.testing.assert_close(ingest.relative_poses(poses[0]).cpu(), torch.from_numpy(ref["f2/poses_rel"][0]),
                               rtol=1e-5, atol=1e-5)
