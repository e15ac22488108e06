"""Dataset output contract -> device ingest.

gradslam's loaders (gradslam/datasets/icl.py:393-533, tum.py, scannet.py) read 8-bit colour and 16-bit depth from disk,
convert them to float32 on the HOST (colour as float(u8), optionally / 255; depth as u16 / scaling_factor) and only then
ship 16 bytes per pixel to the device.  Here the raw 5 bytes per pixel are uploaded (pinned memory, asynchronous) and the
same conversion runs on the device (csrc/gsx_ingest.cu), bit-identical to the host-side one.  The image resize the
loaders can also do is not covered: pass frames at their final size (and `scale_intrinsics` if they were resized).
"""
from typing import Optional, Union

import torch

from . import _C
from .structures.rgbdimages import RGBDImages

__all__ = ["scale_intrinsics", "raw_to_float", "rgbdimages_from_raw", "RawRGBD"]


def scale_intrinsics(intrinsics: torch.Tensor, h_ratio: Union[float, int], w_ratio: Union[float, int]) -> torch.Tensor:
    """Intrinsics of frames resized by (h_ratio, w_ratio) (mirror of gradslam/datasets/datautils.py:73-122)."""
    if not torch.is_tensor(intrinsics):
        raise TypeError("Unsupported input intrinsics type {}".format(type(intrinsics)))
    if not (intrinsics.shape[-2:] == (3, 3) or intrinsics.shape[-2:] == (4, 4)):
        raise ValueError("intrinsics must have shape (*, 3, 3) or (*, 4, 4), but had shape {} instead".format(
            intrinsics.shape))
    out = intrinsics.to(torch.float).clone()
    out[..., 0, 0] *= w_ratio
    out[..., 1, 1] *= h_ratio
    out[..., 0, 2] *= w_ratio
    out[..., 1, 2] *= h_ratio
    return out


def _check_raw(colors, depths):
    if not (torch.is_tensor(colors) and colors.dtype == torch.uint8 and colors.shape[-1] == 3):
        raise TypeError("colors must be a uint8 tensor (..., H, W, 3)")
    if not (torch.is_tensor(depths) and depths.dtype in (torch.uint16, torch.int16)):
        raise TypeError("depths must be a uint16 tensor (..., H, W) or (..., H, W, 1) (int16 storage of the same bits is accepted)")
    if depths.dim() == colors.dim() and depths.shape[-1] == 1:
        depths = depths[..., 0]
    if tuple(depths.shape) != tuple(colors.shape[:-1]):
        raise ValueError("colors {} and depths {} do not describe the same frames".format(tuple(colors.shape), tuple(depths.shape)))
    return colors, depths


def raw_to_float(colors_u8: torch.Tensor, depths_u16: torch.Tensor, scaling_factor: float = 5000.0,
                 normalize_color: bool = False, out_rgb: Optional[torch.Tensor] = None,
                 out_depth: Optional[torch.Tensor] = None):
    """Device-side conversion of CUDA uint8 colour (...,H,W,3) / uint16 depth (...,H,W) to float32 (…,3) / (…,1)."""
    colors_u8, depths_u16 = _check_raw(colors_u8, depths_u16)
    if not colors_u8.is_cuda or not depths_u16.is_cuda:
        raise RuntimeError("gradslam_b200: raw_to_float needs CUDA tensors; there is no CPU path")
    colors_u8, depths_u16 = colors_u8.contiguous(), depths_u16.contiguous()
    dev = colors_u8.device
    rgb = out_rgb if out_rgb is not None else torch.empty(colors_u8.shape, dtype=torch.float32, device=dev)
    depth = out_depth if out_depth is not None else torch.empty((*depths_u16.shape, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _C.lib().gsx_ingest_raw(_C.ptr(colors_u8), _C.ptr(depths_u16), depths_u16.numel(), float(scaling_factor),
                                     1 if normalize_color else 0, _C.ptr(rgb), _C.ptr(depth), _C.stream_ptr(dev))
    _C.check(rc, "gsx_ingest_raw")
    return rgb, depth


class RawRGBD(object):
    """A (B, L) batch of sequences in dataset-native form: uint8 colour (B,L,H,W,3), uint16 depth (B,L,H,W), float32
    intrinsics (B,1,4,4) and poses (B,L,4,4), on the host (pin the image tensors) or on the device.
    `PointFusion(odom='gt')(raw)` uploads and converts chunk by chunk, overlapped with the fusion."""

    def __init__(self, colors_u8, depths_u16, intrinsics, poses, scaling_factor: float = 5000.0,
                 normalize_color: bool = False):
        self.colors, self.depths = _check_raw(colors_u8, depths_u16)
        if self.colors.dim() != 5:
            raise ValueError("colors must have shape (B, L, H, W, 3)")
        self.intrinsics, self.poses = intrinsics, poses
        self.scaling_factor, self.normalize_color = float(scaling_factor), bool(normalize_color)
        self.shape = tuple(self.colors.shape[:4])


def rgbdimages_from_raw(colors_u8, depths_u16, intrinsics, poses=None, *, scaling_factor: float = 5000.0,
                        normalize_color: bool = False, device: Union[torch.device, str] = "cuda") -> RGBDImages:
    """Uploads raw frames (B,L,H,W,3) uint8 / (B,L,H,W) uint16 and returns the float32 RGBDImages the reference's loaders
    would have produced (same bits), resident on `device`."""
    colors_u8, depths_u16 = _check_raw(colors_u8, depths_u16)
    dev = torch.device(device)
    rgb, depth = raw_to_float(colors_u8.to(dev, non_blocking=True), depths_u16.to(dev, non_blocking=True),
                              scaling_factor, normalize_color)
    return RGBDImages(rgb, depth, intrinsics.to(dev), None if poses is None else poses.to(dev))
