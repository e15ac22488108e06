"""Backward of K1 (depth -> vertex / normal / global maps) against PyTorch autograd of a plain fp32 torch
implementation of the same op chain (gradslam/structures/rgbdimages.py:643-762)."""
import pytest
import torch

from gradslam_b200.synthetic import make_sequence

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _torch_maps(depth, K, poses):
    """Differentiable fp32 torch reference (einsum-free, same formulas).  depth (B,L,H,W,1)."""
    B, L, H, W, _ = depth.shape
    fx, fy, cx, cy = K[:, 0, 0, 0] + 1e-6, K[:, 0, 1, 1] + 1e-6, K[:, 0, 0, 2], K[:, 0, 1, 2]
    u = torch.arange(W, dtype=torch.float32, device=depth.device).view(1, 1, 1, W)
    v = torch.arange(H, dtype=torch.float32, device=depth.device).view(1, 1, H, 1)
    d = depth[..., 0]
    vf = (d > 0).float()
    x = ((u - cx.view(B, 1, 1, 1)) / fx.view(B, 1, 1, 1)) * d * vf
    y = ((v - cy.view(B, 1, 1, 1)) / fy.view(B, 1, 1, 1)) * d * vf
    vert = torch.stack([x.expand(B, L, H, W), y.expand(B, L, H, W), d * vf], -1)
    dh = torch.zeros_like(vert)
    dv = torch.zeros_like(vert)
    dh[..., :, :-1, :] = vert[..., :, 1:, :] - vert[..., :, :-1, :]
    dv[..., :-1, :, :] = vert[..., 1:, :, :] - vert[..., :-1, :, :]
    dh = torch.cat([dh[..., :, :-1, :], dh[..., :, -2:-1, :]], dim=-2)
    dv = torch.cat([dv[..., :-1, :, :], dv[..., -2:-1, :, :]], dim=-3)
    c = torch.cross(dh, dv, dim=-1)
    nrm = torch.linalg.norm(c, dim=-1, keepdim=True)
    n = c / torch.where(nrm == 0, torch.ones_like(nrm), nrm) * vf.unsqueeze(-1)
    R, t = poses[..., :3, :3], poses[..., :3, 3]
    gv = (torch.einsum("blij,blhwj->blhwi", R, vert) + t.view(B, L, 1, 1, 3)) * vf.unsqueeze(-1)
    gn = torch.einsum("blij,blhwj->blhwi", R, n)
    return vert, n, gv, gn


@pytest.mark.parametrize("shape", [(2, 2, 24, 40), (1, 1, 17, 23)])
def test_backproject_backward_matches_autograd(shape):
    import gradslam_b200 as gs

    B, L, H, W = shape
    rgb, depth, K, poses = make_sequence(B, L, H, W, seed=21, isolated_holes=True)  # no degenerate cross products
    g = torch.Generator().manual_seed(3)
    ups = [torch.randn(B, L, H, W, 3, generator=g).to(DEV) for _ in range(4)]
    # engine
    d1 = depth.to(DEV).requires_grad_(True)
    p1 = poses.to(DEV).requires_grad_(True)
    fr = gs.RGBDImages(rgb.to(DEV), d1, K.to(DEV), p1)
    outs = (fr.vertex_map, fr.normal_map, fr.global_vertex_map, fr.global_normal_map)
    loss = sum((o * w).sum() for o, w in zip(outs, ups))
    loss.backward()
    # torch reference
    d2 = depth.to(DEV).requires_grad_(True)
    p2 = poses.to(DEV).requires_grad_(True)
    refs = _torch_maps(d2, K.to(DEV), p2)
    for o, r in zip(outs, refs):
        torch.testing.assert_close(o.detach(), r.detach(), rtol=1e-4, atol=1e-5)
    sum((o * w).sum() for o, w in zip(refs, ups)).backward()
    scale = d2.grad.abs().max().item()
    torch.testing.assert_close(d1.grad, d2.grad, rtol=1e-3, atol=1e-4 * scale)
    pscale = p2.grad.abs().max().item()
    torch.testing.assert_close(p1.grad[..., :3, :], p2.grad[..., :3, :], rtol=1e-3, atol=1e-4 * pscale)
    assert p1.grad[..., 3, :].abs().max() == 0


def test_backward_only_global_maps_and_no_pose_grad():
    import gradslam_b200 as gs

    rgb, depth, K, poses = make_sequence(1, 2, 20, 28, seed=22, isolated_holes=True)
    d1 = depth.to(DEV).requires_grad_(True)
    fr = gs.RGBDImages(rgb.to(DEV), d1, K.to(DEV), poses.to(DEV))
    w = torch.randn(1, 2, 20, 28, 3, device=DEV)
    (fr.global_vertex_map * w).sum().backward()
    d2 = depth.to(DEV).requires_grad_(True)
    refs = _torch_maps(d2, K.to(DEV), poses.to(DEV))
    (refs[2] * w).sum().backward()
    torch.testing.assert_close(d1.grad, d2.grad, rtol=1e-3, atol=1e-5)
