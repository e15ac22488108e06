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
from .fusionutils import update_map_fusion
from .icpslam import ICPSLAM

__all__ = ["PointFusion"]


class _SequenceWorkspace:
    """Scratch of gsx_pointfusion_sequence_gt (two frame workspaces used alternately), per (device, stream, B, H, W)."""

    _cache = {}

    def __init__(self, device, B, H, W):
        self.buf = torch.zeros(_C.lib().gsx_pointfusion_sequence_workspace_bytes(B, H, W), dtype=torch.uint8,
                               device=device)

    @classmethod
    def get(cls, device, B, H, W):
        key = (str(device), torch.cuda.current_stream(device).cuda_stream, B, H, W)
        if key not in cls._cache:
            cls._cache[key] = cls(device, B, H, W)
        return cls._cache[key]


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

    def forward(self, frames, out=None):
        """As ICPSLAM.forward; additionally accepts a `gradslam_b200.ingest.RawRGBD` batch (uint8 colour + uint16 depth)
        when odom='gt': the raw frames are uploaded and converted on the device, overlapped with the fusion."""
        from ..ingest import RawRGBD

        if isinstance(frames, RawRGBD):
            if self.odom != "gt" or frames.poses is None:
                raise ValueError("RawRGBD input is supported for odom='gt' with poses; convert with "
                                 "ingest.rgbdimages_from_raw for the other odometry modes")
            self._check_out(out, frames.shape[0])
            return self._forward_sequence(frames, raw=True, out=out)
        return super().forward(frames, out)

    def _forward_sequence(self, frames, chunk: int = 4, raw: bool = False, out=None):
        """odom='gt': the whole (B, L) sequence runs as C calls chaining K1 -> K2/K3 -> K4 per frame with no
        host synchronisation.  Frames that live in HOST memory are uploaded `chunk` frames at a time on a side
        stream, so the copy of chunk i+1 overlaps the fusion of chunk i (pin the host tensors for this)."""
        if self.odom != "gt" or frames.poses is None or torch.is_tensor(self.sigma):
            return None
        dev = self.device
        if raw:
            B, L, H, W = frames.shape
            src_depth, src_rgb = frames.depths, frames.colors
            if src_depth.device == dev:  # already uploaded: convert in one go
                from ..ingest import raw_to_float

                rgb_f, depth_f = raw_to_float(src_rgb, src_depth, frames.scaling_factor, frames.normalize_color)
                return self._forward_sequence(RGBDImages(rgb_f, depth_f, frames.intrinsics.to(dev),
                                                         frames.poses.to(dev)), out=out)
            on_device = False
        else:
            if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (
                    frames.depth_image, frames.rgb_image, frames.poses, frames.intrinsics)):
                return None  # differentiable mode goes through the per-frame step loop
            if frames.channels_first:
                frames = frames.to_channels_last()
            B, L, H, W = frames.shape
            src_depth, src_rgb = frames.depth_image, frames.rgb_image
            on_device = src_depth.device == dev
        P = H * W
        K = frames.intrinsics.to(dev).contiguous()
        poses = frames.poses.to(dev).contiguous()
        if on_device:
            depth, rgb = src_depth.contiguous(), src_rgb.contiguous()
            chunk = L
        else:
            depth = torch.empty((B, L, H, W, 1), dtype=torch.float32, device=dev)
            rgb = torch.empty((B, L, H, W, 3), dtype=torch.float32, device=dev)
            copy_stream = torch.cuda.Stream(device=dev)
            if raw:  # staging buffers for the 5-byte pixels; converted chunk by chunk on the compute stream
                raw_depth = torch.empty((B, L, H, W), dtype=src_depth.dtype, device=dev)
                raw_rgb = torch.empty((B, L, H, W, 3), dtype=torch.uint8, device=dev)
        for t, name in ((depth, "depth_image"), (rgb, "rgb_image"), (K, "intrinsics"), (poses, "poses")):
            _C.require_cuda(t, name)
        if out is not None:  # caller-provided storage (validated by forward: empty, B maps, this device)
            if not (out.has_normals and out.has_colors and out._has_cc):
                raise ValueError("out must have normals, colours and a confidence count per point for map fusion")
            pc = out
            pc._cur = 0
        else:
            pc = Pointclouds(device=dev)
            pc._allocate(B, L * P, 1, zero=False)
        ws = _SequenceWorkspace.get(dev, B, H, W)
        main = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            ready = []
            if not on_device:
                copy_stream.wait_stream(main)  # the fresh buffers must exist before the copies start
                with torch.cuda.stream(copy_stream):
                    for s0 in range(0, L, chunk):
                        s1 = min(L, s0 + chunk)
                        dst_d, dst_c = (raw_depth, raw_rgb) if raw else (depth, rgb)
                        for b in range(B):  # per-element slices are contiguous: true async DMA from pinned memory
                            dst_d[b, s0:s1].copy_(src_depth[b, s0:s1], non_blocking=True)
                            dst_c[b, s0:s1].copy_(src_rgb[b, s0:s1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        ready.append(ev)
            for i, s0 in enumerate(range(0, L, chunk)):
                s1 = min(L, s0 + chunk)
                if ready:
                    main.wait_event(ready[i])
                if raw:  # u8 / u16 -> float32 for this chunk (per element: the chunk is contiguous inside an element)
                    for b in range(B):
                        _C.check(_C.lib().gsx_ingest_raw(
                            _C.ptr(raw_rgb[b, s0:s1]), _C.ptr(raw_depth[b, s0:s1]), (s1 - s0) * P,
                            frames.scaling_factor, 1 if frames.normalize_color else 0, _C.ptr(rgb[b, s0:s1]),
                            _C.ptr(depth[b, s0:s1]), _C.stream_ptr(dev)), "gsx_ingest_raw")
                rc = _C.lib().gsx_pointfusion_sequence_gt(
                    _C.ptr(pc._geo), _C.ptr(pc._col), _C.ptr(pc._counts_dev), pc.capacity, min(s0 * P, pc.capacity),
                    _C.ptr(depth), _C.ptr(rgb), _C.ptr(K), _C.ptr(poses), B, L, s0, s1, H, W, float(self.dist_th),
                    float(self.dot_th), float(self.sigma), _C.ptr(ws.buf), _C.ptr(pc._overflow_flag()),
                    _C.stream_ptr(dev))
                _C.check(rc, "gsx_pointfusion_sequence_gt")
            if not on_device:
                for t in ((raw_depth, raw_rgb) if raw else (depth, rgb)):
                    t.record_stream(copy_stream)
        pc._cur = L & 1
        pc._counts_host = None
        pc._bound = pc.capacity
        pc._list_cache = {}
        pc._tail_dirty = pc._uninit
        return pc, poses.clone()
