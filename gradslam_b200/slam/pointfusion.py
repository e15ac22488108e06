"""PointFusion: ICPSLAM whose mapping step is confidence-weighted surfel fusion.

Host-side mirror of gradslam.slam.PointFusion (gradslam/slam/pointfusion.py:16-112): same keywords and
defaults (dist_th=0.05, angle_th=20, sigma=0.6), `dot_th = cos(angle_th)`.
"""
import math
import warnings
from typing import Union

import torch

from .. import _C
from ..structures.pointclouds import Pointclouds
from ..structures.rgbdimages import RGBDImages
from .fusionutils import _Workspace, update_map_fusion
from .icpslam import ICPSLAM

__all__ = ["PointFusion"]


class PointFusion(ICPSLAM):
    def __init__(self, *, odom: str = "gradicp", dist_th: Union[float, int] = 0.05,
                 angle_th: Union[float, int] = 20, sigma: Union[float, int] = 0.6, dsratio: int = 4,
                 numiters: int = 20, damp: float = 1e-8, dist_thresh: Union[float, int, None] = None,
                 lambda_max: Union[float, int] = 2.0, B: Union[float, int] = 1.0, B2: Union[float, int] = 1.0,
                 nu: Union[float, int] = 200.0, device: Union[torch.device, str, None] = None):
        super().__init__(odom=odom, dsratio=dsratio, numiters=numiters, damp=damp, dist_thresh=dist_thresh,
                         lambda_max=lambda_max, B=B, B2=B2, nu=nu, device=device)
        if not isinstance(dist_th, (float, int)):
            raise TypeError("Distance threshold must be of type float or int; but was of type {}.".format(
                type(dist_th)))
        if not isinstance(angle_th, (float, int)):
            raise TypeError("Angle threshold must be of type float or int; but was of type {}.".format(
                type(angle_th)))
        if dist_th < 0:
            warnings.warn("Distance threshold ({}) should be non-negative.".format(dist_th))
        if not ((0 <= angle_th) and (angle_th <= 90)):
            warnings.warn("Angle threshold ({}) should be non-negative and <=90.".format(angle_th))
        self.dist_th = dist_th
        self.dot_th = math.cos((angle_th * math.pi) / 180)
        self.sigma = sigma

    def _map(self, pointclouds: Pointclouds, live_frame: RGBDImages, inplace: bool = False):
        return update_map_fusion(pointclouds, live_frame, self.dist_th, self.dot_th, self.sigma, inplace)

    def _forward_sequence(self, frames: RGBDImages):
        """odom='gt': the whole (B, L) sequence is ONE C call chaining K1 -> K2/K3 -> K4 per frame."""
        if self.odom != "gt" or frames.poses is None or torch.is_tensor(self.sigma):
            return None
        frames = frames.to(self.device).to_channels_last()
        B, L, H, W = frames.shape
        depth, rgb = frames.depth_image.contiguous(), frames.rgb_image.contiguous()
        _C.require_cuda(depth, "depth_image")
        K, poses = frames.intrinsics.contiguous(), frames.poses.contiguous()
        P = H * W
        pc = Pointclouds(device=self.device)
        pc._allocate(B, L * P, 1)
        ws = _Workspace.get(self.device, B, H, W)
        scratch = torch.empty((2, B, H, W, 3), dtype=torch.float32, device=self.device)
        st = pc._store
        with torch.cuda.device(self.device):
            rc = _C.lib().gsx_pointfusion_sequence_gt(
                _C.ptr(st["points"]), _C.ptr(st["normals"]), _C.ptr(st["colors"]), _C.ptr(st["features"]),
                _C.ptr(pc._counts_dev), pc.capacity, 0, _C.ptr(depth), _C.ptr(rgb), _C.ptr(K), _C.ptr(poses),
                B, L, H, W, float(self.dist_th), float(self.dot_th), float(self.sigma), _C.ptr(scratch),
                _C.ptr(ws.buf), ws.next_epochs(L), _C.ptr(pc._overflow_flag()), _C.stream_ptr(self.device))
        _C.check(rc, "gsx_pointfusion_sequence_gt")
        if L & 1:
            pc._cur ^= 1
        pc._counts_host = None
        pc._bound = pc.capacity
        pc._list_cache = {}
        return pc, poses.clone()
