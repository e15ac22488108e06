"""The three geometryutils helpers the hot path touches (mirror of gradslam/geometry/geometryutils.py:
create_meshgrid :576, transform_pointcloud :737, relative_transformation :413) plus the two rigid 4x4
helpers gradslam imports from kornia.geometry.linalg (compose_transformations / inverse_transformation;
call sites slam/icpslam.py:245, slam/fusionutils.py:249).  4x4 plumbing only.
"""
from typing import Optional

import torch

__all__ = ["create_meshgrid", "transform_pointcloud", "relative_transformation", "compose_transformations",
           "inverse_transformation"]


def compose_transformations(trans_01: torch.Tensor, trans_12: torch.Tensor) -> torch.Tensor:
    """Rigid product T01 · T12 with the bottom row forced to [0, 0, 0, 1]."""
    if not (torch.is_tensor(trans_01) and torch.is_tensor(trans_12)):
        raise TypeError("Inputs must be torch.Tensor")
    R = trans_01[..., :3, :3] @ trans_12[..., :3, :3]
    t = trans_01[..., :3, :3] @ trans_12[..., :3, 3:] + trans_01[..., :3, 3:]
    out = torch.zeros_like(trans_01)
    out[..., :3, :3] = R
    out[..., :3, 3:] = t
    out[..., 3, 3] = 1.0
    return out


def inverse_transformation(trans_12: torch.Tensor) -> torch.Tensor:
    """[R^T, -R^T t] with the bottom row forced to [0, 0, 0, 1]."""
    if not torch.is_tensor(trans_12):
        raise TypeError("Input must be torch.Tensor")
    Rt = trans_12[..., :3, :3].transpose(-1, -2)
    out = torch.zeros_like(trans_12)
    out[..., :3, :3] = Rt
    out[..., :3, 3:] = (-Rt) @ trans_12[..., :3, 3:]
    out[..., 3, 3] = 1.0
    return out


def relative_transformation(trans_01: torch.Tensor, trans_02: torch.Tensor,
                            orthogonal_rotations: bool = False) -> torch.Tensor:
    """T12 = T01^{-1} · T02 for (N,4,4) or (4,4) inputs."""
    if not torch.is_tensor(trans_01):
        raise TypeError("Input trans_01 type is not a torch.Tensor. Got {}".format(type(trans_01)))
    if not torch.is_tensor(trans_02):
        raise TypeError("Input trans_02 type is not a torch.Tensor. Got {}".format(type(trans_02)))
    if not trans_01.dim() in (2, 3) and trans_01.shape[-2:] == (4, 4):
        raise ValueError("Input must be a of the shape Nx4x4 or 4x4. Got {}".format(trans_01.shape))
    if not trans_02.dim() in (2, 3) and trans_02.shape[-2:] == (4, 4):
        raise ValueError("Input must be a of the shape Nx4x4 or 4x4. Got {}".format(trans_02.shape))
    if not trans_01.dim() == trans_02.dim():
        raise ValueError("Input number of dims must match. Got {} and {}".format(trans_01.dim(), trans_02.dim()))
    trans_10 = inverse_transformation(trans_01) if orthogonal_rotations else torch.inverse(trans_01)
    return compose_transformations(trans_10, trans_02)


def create_meshgrid(height: int, width: int, normalized_coords: Optional[bool] = True) -> torch.Tensor:
    """(1, H, W, 2) grid holding (row, col) coordinates, optionally normalised to [-1, 1]."""
    if normalized_coords:
        rows, cols = torch.linspace(-1, 1, height), torch.linspace(-1, 1, width)
    else:
        rows, cols = torch.linspace(0, height - 1, height), torch.linspace(0, width - 1, width)
    grid = torch.stack(torch.meshgrid(rows, cols, indexing="ij"), dim=-1)
    return grid.unsqueeze(0)


def transform_pointcloud(pointcloud: torch.Tensor, transform: torch.Tensor) -> torch.Tensor:
    """Applies a 4x4 rigid transform to an (N, 3) cloud."""
    if not torch.is_tensor(pointcloud):
        raise TypeError("pointcloud should be tensor, but was %r instead" % type(pointcloud))
    if not torch.is_tensor(transform):
        raise TypeError("transform should be tensor, but was %r instead" % type(transform))
    if not pointcloud.ndim == 2:
        raise ValueError("pointcloud should have ndim of 2, but had {} instead.".format(pointcloud.ndim))
    if not pointcloud.shape[1] == 3:
        raise ValueError("pointcloud.shape[1] should be 3 (x, y, z), but was {} instead.".format(pointcloud.shape[1]))
    if not transform.shape[-2:] == (4, 4):
        raise ValueError("transform should be of shape (4, 4), but was {} instead.".format(transform.shape))
    return pointcloud @ transform[:3, :3].transpose(0, 1) + transform[:3, 3]
