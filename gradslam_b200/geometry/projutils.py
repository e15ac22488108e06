"""Projective helpers used on the hot path's edges (mirror of gradslam/geometry/projutils.py).

Only small tensor plumbing lives here; the per-pixel / per-point projection arithmetic of the hot path is
inside the CUDA kernels (csrc/gsx_frame.cu, csrc/gsx_fusion.cu).
"""
import torch

__all__ = ["homogenize_points", "unhomogenize_points", "project_points", "unproject_points", "inverse_intrinsics"]


def _need_tensor(x, name):
    if not torch.is_tensor(x):
        raise TypeError("Expected input {} to be of type torch.Tensor. Got {} instead.".format(name, type(x)))


def homogenize_points(pts: torch.Tensor) -> torch.Tensor:
    """(*, D) -> (*, D+1) by appending ones (projutils.py:10-43)."""
    _need_tensor(pts, "pts")
    if pts.dim() < 2:
        raise ValueError("Input tensor must have at least 2 dimensions. Got {} instad.".format(pts.dim()))
    return torch.nn.functional.pad(pts, (0, 1), "constant", 1.0)


def unhomogenize_points(pts: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """(*, D) -> (*, D-1): divide by the last coordinate, using 1 where |w| <= eps (projutils.py:46-89)."""
    _need_tensor(pts, "pts")
    if pts.dim() < 2:
        raise ValueError("Input tensor must have at least 2 dimensions. Got {} instad.".format(pts.dim()))
    w = pts[..., -1:]
    scale = torch.where(torch.abs(w) > eps, 1.0 / w, torch.ones_like(w))
    return scale * pts[..., :-1]


def project_points(cam_coords: torch.Tensor, proj_mat: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Pinhole projection (*, 3|4) x (*, 4, 4) -> (*, 2); z == 0 divides by 1 (projutils.py:92-238)."""
    _need_tensor(cam_coords, "cam_coords")
    _need_tensor(proj_mat, "proj_mat")
    if cam_coords.dim() < 2:
        raise ValueError("Input cam_coords must have at least 2 dims. Got {} instead.".format(cam_coords.dim()))
    if cam_coords.shape[-1] not in (3, 4):
        raise ValueError("Input cam_coords must have shape (*, 3), or (*, 4). Got {} instead.".format(cam_coords.shape))
    if proj_mat.dim() < 2:
        raise ValueError("Input proj_mat must have at least 2 dims. Got {} instead.".format(proj_mat.dim()))
    if proj_mat.shape[-1] != 4 or proj_mat.shape[-2] != 4:
        raise ValueError("Input proj_mat must have shape (*, 4, 4). Got {} instead.".format(proj_mat.shape))
    if proj_mat.dim() > 2 and proj_mat.dim() != cam_coords.dim():
        raise ValueError("Input proj_mat must either have 2 dimensions, or have equal number of dimensions to "
                         "cam_coords. Got {} instead.".format(proj_mat.dim()))
    if proj_mat.dim() > 2 and proj_mat.shape[0] != cam_coords.shape[0]:
        raise ValueError("Batch sizes of proj_mat and cam_coords do not match. Shapes: {} and {} respectively.".format(
            proj_mat.shape, cam_coords.shape))
    homo = homogenize_points(cam_coords) if cam_coords.shape[-1] == 3 else cam_coords
    if proj_mat.dim() == 2:
        out = homo @ proj_mat.transpose(0, 1)
    else:
        out = torch.matmul(proj_mat.unsqueeze(-3), homo.unsqueeze(-1)).squeeze(-1)
    z = out[..., 2]
    den = torch.where(z != 0, z, torch.ones_like(z))
    return torch.stack((out[..., 0] / den, out[..., 1] / den), dim=-1)


def unproject_points(pixel_coords: torch.Tensor, intrinsics_inv: torch.Tensor, depths: torch.Tensor) -> torch.Tensor:
    """Back-projects (*, 2|3) pixels with depths (*,) or (*, 1) -> camera points (*, 3) (projutils.py:241-402)."""
    _need_tensor(pixel_coords, "pixel_coords")
    _need_tensor(intrinsics_inv, "intrinsics_inv")
    _need_tensor(depths, "depths")
    if pixel_coords.dim() < 2:
        raise ValueError("Input pixel_coords must have at least 2 dims. Got {} instead.".format(pixel_coords.dim()))
    if pixel_coords.shape[-1] not in (2, 3):
        raise ValueError("Input pixel_coords must have shape (*, 2), or (*, 3). Got {} instead.".format(
            pixel_coords.shape))
    if intrinsics_inv.dim() < 2:
        raise ValueError("Input intrinsics_inv must have at least 2 dims. Got {} instead.".format(intrinsics_inv.dim()))
    if intrinsics_inv.shape[-1] != 3 or intrinsics_inv.shape[-2] != 3:
        raise ValueError("Input intrinsics_inv must have shape (*, 3, 3). Got {} instead.".format(intrinsics_inv.shape))
    if intrinsics_inv.dim() > 2 and intrinsics_inv.dim() != pixel_coords.dim():
        raise ValueError("Input intrinsics_inv must either have 2 dimensions, or have equal number of dimensions to "
                         "pixel_coords. Got {} instead.".format(intrinsics_inv.dim()))
    if intrinsics_inv.dim() > 2 and intrinsics_inv.shape[0] != pixel_coords.shape[0]:
        raise ValueError("Batch sizes of intrinsics_inv and pixel_coords do not match. Shapes: {} and {} "
                         "respectively.".format(intrinsics_inv.shape, pixel_coords.shape))
    if pixel_coords.shape[:-1] != depths.shape:
        if depths.shape[-1] != 1 or pixel_coords.shape[:-1] != depths.shape[:-1]:
            raise ValueError("Input pixel_coords and depths must have the same shape for all dimensions except "
                             "the last.  Got {} and {} respectively.".format(pixel_coords.shape, depths.shape))
    homo = homogenize_points(pixel_coords) if pixel_coords.shape[-1] == 2 else pixel_coords
    if depths.dim() < homo.dim():
        depths = depths.unsqueeze(-1)
    if intrinsics_inv.dim() == 2:
        rays = homo @ intrinsics_inv.transpose(0, 1)
    else:
        rays = torch.matmul(intrinsics_inv.unsqueeze(-3), homo.unsqueeze(-1)).squeeze(-1)
    return rays * depths


def inverse_intrinsics(K: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Closed-form inverse of a pinhole K (3x3 or 4x4); eps is added to fx, fy (projutils.py:405-450)."""
    if not torch.is_tensor(K):
        raise TypeError("Expected K to be of type torch.Tensor. Got {0} instead.".format(type(K)))
    if K.dim() < 2:
        raise ValueError("Input K must have at least 2 dims. Got {0} instead.".format(K.dim()))
    if not ((K.shape[-1] == 3 and K.shape[-2] == 3) or (K.shape[-1] == 4 and K.shape[-2] == 4)):
        raise ValueError("Input K must have shape (*, 4, 4) or (*, 3, 3). Got {0} instead.".format(K.shape))
    fx, fy = K[..., 0, 0] + eps, K[..., 1, 1] + eps
    Kinv = torch.zeros_like(K)
    Kinv[..., 0, 0] = 1.0 / fx
    Kinv[..., 1, 1] = 1.0 / fy
    Kinv[..., 0, 2] = -1.0 * K[..., 0, 2] / fx
    Kinv[..., 1, 2] = -1.0 * K[..., 1, 2] / fy
    Kinv[..., 2, 2] = 1
    Kinv[..., -1, -1] = 1
    return Kinv
