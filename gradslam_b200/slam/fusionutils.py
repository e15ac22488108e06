"""PointFusion map update: projective data association + confidence-weighted surfel fusion.

Host-side mirror of gradslam/slam/fusionutils.py (same function names, arguments, return types, errors and
warnings).  `update_map_fusion` runs as two hand-written sm_100a kernels over an in-place, capacity-backed
map (csrc/gsx_fusion.cu): no table is materialised, no whole-map clone / cat per frame, no host sync.  The
table-returning helpers (`find_active_map_points`, `find_similar_map_points`,
`find_best_unique_correspondences`, `fuse_with_map`) are kept for API parity and run the same arithmetic
through the table kernels in csrc/gsx_tables.cu.
"""
import warnings
from typing import Union

import torch

from .. import _C
from ..structures.pointclouds import Pointclouds
from ..structures.rgbdimages import RGBDImages, _frame_base

__all__ = ["update_map_fusion", "update_map_aggregate"]


# --------------------------------------------------------------------------------------------- small helpers
def get_alpha(points: torch.Tensor, sigma: Union[torch.Tensor, float, int], dim: int = -1, keepdim: bool = False,
              eps: float = 1e-7) -> torch.Tensor:
    """Sample confidence alpha = clamp(exp(-|p|^2 / 2 sigma^2), eps, 1.01) (fusionutils.py:16-73).
    Differentiable torch helper; inside the fused update alpha is evaluated by the merge kernel."""
    if not torch.is_tensor(points):
        raise TypeError("Expected input points to be of type torch.Tensor. Got {0} instead.".format(type(points)))
    if not (torch.is_tensor(sigma) or isinstance(sigma, (float, int))):
        raise TypeError("Expected input sigma to be of type torch.Tensor or float or int. Got {0} instead.".format(
            type(sigma)))
    if not isinstance(eps, float):
        raise TypeError("Expected input eps to be of type float. Got {0} instead.".format(type(eps)))
    if points.shape[dim] != 3:
        raise ValueError("Expected length of dim-th ({0}th) dimension to be 3. Got {1} instead.".format(
            dim, points.shape[dim]))
    if torch.is_tensor(sigma) and sigma.ndim != 0:
        raise ValueError("Expected sigma.ndim to be 0 (scalar). Got {0}.".format(sigma.ndim))
    sq = torch.sum(points ** 2, dim, keepdim=keepdim)
    return torch.clamp(torch.exp(-sq / (2 * (sigma ** 2))), min=eps, max=1.01)


def _pair_checks(tensor1, tensor2, th, th_name, dim):
    for name, t in (("tensor1", tensor1), ("tensor2", tensor2)):
        if not torch.is_tensor(t):
            raise TypeError("Expected input {} to be of type torch.Tensor. Got {} instead.".format(name, type(t)))
    if not isinstance(th, (float, int)):
        raise TypeError("Expected input {} to be of type float or int. Got {} instead.".format(th_name, type(th)))
    if tensor1.shape != tensor2.shape:
        raise ValueError("tensor1 and tensor2 should have the same shape, but had shapes {0} and {1} "
                         "respectively.".format(tensor1.shape, tensor2.shape))
    if tensor1.shape[dim] != 3:
        raise ValueError("Expected length of input tensors' dim-th ({0}th) dimension to be 3. Got {1} "
                         "instead.".format(dim, tensor1.shape[dim]))


def are_points_close(tensor1: torch.Tensor, tensor2: torch.Tensor, dist_th: Union[float, int],
                     dim: int = -1) -> torch.Tensor:
    """||t1 - t2|| < dist_th along `dim` (fusionutils.py:76-130)."""
    _pair_checks(tensor1, tensor2, dist_th, "dist_th", dim)
    return (tensor1 - tensor2).norm(dim=dim) < dist_th


def are_normals_similar(tensor1: torch.Tensor, tensor2: torch.Tensor, dot_th: Union[float, int],
                        dim: int = -1) -> torch.Tensor:
    """<t1, t2> > dot_th along `dim`; warns if the inputs were not unit length (fusionutils.py:133-195)."""
    _pair_checks(tensor1, tensor2, dot_th, "dot_th", dim)
    dots = (tensor1 * tensor2).sum(dim)
    if dots.numel() > 0 and dots.max() > 1.001:
        warnings.warn("Max of dot product was {0} > 1. Inputs were not normalized along dim ({1}). Was this "
                      "intentional?".format(dots.max(), dim), RuntimeWarning)
    return dots > dot_th


# --------------------------------------------------------------------------------------------- workspaces
class _Workspace:
    """Per (device, B, H, W) scratch for the fusion kernels: arg-min records + scan state, zeroed once."""

    _cache = {}

    def __init__(self, device, B, H, W):
        nbytes = _C.lib().gsx_fusion_workspace_bytes(B, H, W)
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.epoch = 0

    @classmethod
    def get(cls, device, B, H, W):
        key = (str(device), B, H, W)
        ws = cls._cache.get(key)
        if ws is None:
            ws = cls(device, B, H, W)
            cls._cache[key] = ws
        return ws

    def next_epochs(self, n=1):
        first = self.epoch + 1
        self.epoch += n
        if self.epoch >= (1 << 30) - 1:  # wrap: start over on a clean buffer
            self.buf.zero_()
            self.epoch = n
            first = 1
        return first


def _check_frame(rgbdimages):
    if not isinstance(rgbdimages, RGBDImages):
        raise TypeError("Expected rgbdimages to be of type gradslam.RGBDImages. Got {0}.".format(type(rgbdimages)))
    if rgbdimages.shape[1] != 1:
        raise ValueError("Expected rgbdimages to have sequence length of 1. Got {0}.".format(rgbdimages.shape[1]))


def _launch_merge_append(pointclouds, frames, vmap, nmap, sigma):
    """K4 on the current workspace state.  vmap/nmap: (B,1,H,W,3) maps to merge/append, or None for both:
    the kernel then samples world-frame vertex/normal from depth on the fly (needs frames.poses)."""
    B, _, H, W = frames.shape
    P = H * W
    dev = pointclouds.device
    ws = _Workspace.get(dev, B, H, W)
    depth, d_bs = _frame_base(frames.depth_image, P)
    rgb, c_bs = _frame_base(frames.rgb_image, P * 3)
    K = frames.intrinsics.contiguous()
    poses = None if vmap is not None else frames.poses.contiguous()
    if vmap is not None:  # cached maps of a sliced sequence are strided views: the kernels want dense (B,H,W,3)
        vmap, nmap = vmap.contiguous(), nmap.contiguous()
    st = pointclouds._store
    cin = pointclouds._counts_dev[pointclouds._cur]
    cout = pointclouds._counts_dev[pointclouds._cur ^ 1]
    with torch.cuda.device(dev):
        rc = _C.lib().gsx_fusion_merge_append(
            _C.ptr(st["points"]), _C.ptr(st["normals"]), _C.ptr(st["colors"]), _C.ptr(st["features"]),
            _C.ptr(cin), _C.ptr(cout), pointclouds.capacity, _C.ptr(depth), d_bs, _C.ptr(rgb), c_bs, _C.ptr(K), 16,
            _C.ptr(poses), 16, _C.ptr(vmap), _C.ptr(nmap), B, H, W, float(sigma), _C.ptr(ws.buf), ws.next_epochs(1),
            _C.ptr(pointclouds._overflow_flag()), _C.stream_ptr(dev))
    _C.check(rc, "gsx_fusion_merge_append")
    pointclouds._mark_device_updated(pointclouds._bound + P)


def _prepare_map(pointclouds, frames, with_features):
    """Makes sure the map has storage for B elements with room for one more frame."""
    B, _, H, W = frames.shape
    if not pointclouds.has_points:
        pointclouds.device = frames.device
        pointclouds._allocate(B, 2 * H * W, 1 if with_features else 0)
    elif len(pointclouds) != B:
        raise ValueError("Expected equal batch sizes for pointclouds and rgbdimages. Got {0} and {1} "
                         "respectively.".format(len(pointclouds), B))
    if pointclouds._bound + H * W > pointclouds.capacity:
        pointclouds._host_counts()  # one sync tightens the bound before we decide to grow
    pointclouds.reserve(pointclouds._bound + H * W)


def _append_valid_pixels(pointclouds, frames, global_coordinates=True, sigma=0.6):
    """Stable append of every valid pixel (K4 with no matches): update_map_aggregate / pointclouds_from_rgbdimages."""
    _check_frame(frames)
    frames = frames.to_channels_last()
    _C.require_cuda(frames.depth_image, "depth_image")
    had_points = pointclouds.has_points
    with_features = pointclouds.has_features if had_points else False
    _prepare_map(pointclouds, frames, with_features)
    if global_coordinates and frames.poses is not None:
        vmap = nmap = None  # sampled on the fly inside the kernel
    else:
        vmap = frames.global_vertex_map if global_coordinates else frames.vertex_map
        nmap = frames.global_normal_map if global_coordinates else frames.normal_map
    _launch_merge_append(pointclouds, frames, vmap, nmap, sigma)
    return pointclouds


def _fused_update(pointclouds, frames, dist_th, dot_th, sigma):
    """K2/K3 then K4, in place."""
    frames = frames.to_channels_last()
    _C.require_cuda(frames.depth_image, "depth_image")
    if frames.poses is None:
        raise ValueError("rgbdimages must have poses for map fusion")
    if pointclouds.has_points:
        for what in ("normals", "colors", "features"):
            if not getattr(pointclouds, "has_" + what):
                raise ValueError("Pointclouds must have {} for map fusion, but did not.".format(what))
        if pointclouds.num_features != 1:
            raise ValueError("Pointclouds features must be a single confidence count per point for map fusion.")
    B, _, H, W = frames.shape
    P = H * W
    _prepare_map(pointclouds, frames, True)
    dev = pointclouds.device
    # Frame geometry: if the caller already materialised the global maps (they are cached on `frames`) use
    # them, otherwise let the kernels sample vertex / normal from depth on the fly (same bits, no K1 launch).
    vmap, nmap = frames._global_vertex_map, frames._global_normal_map
    if vmap is None or nmap is None:
        vmap = nmap = None
    else:
        vmap, nmap = vmap.contiguous(), nmap.contiguous()
    if pointclouds._bound > 0:
        ws = _Workspace.get(dev, B, H, W)
        st = pointclouds._store
        poses, p_bs = frames.poses.contiguous(), 16
        K = frames.intrinsics.contiguous()
        depth, d_bs = _frame_base(frames.depth_image, P)
        with torch.cuda.device(dev):
            rc = _C.lib().gsx_fusion_project_select(
                _C.ptr(st["points"]), _C.ptr(st["normals"]), _C.ptr(st["features"]),
                _C.ptr(pointclouds._counts_dev[pointclouds._cur]), pointclouds.capacity, pointclouds._bound,
                _C.ptr(poses), p_bs, _C.ptr(K), 16, _C.ptr(depth), d_bs, _C.ptr(vmap), _C.ptr(nmap), B, H, W,
                float(dist_th), float(dot_th), _C.ptr(ws.buf), _C.stream_ptr(dev))
        _C.check(rc, "gsx_fusion_project_select")
    _launch_merge_append(pointclouds, frames, vmap, nmap, sigma)
    return pointclouds


# --------------------------------------------------------------------------------------------- table API
def _compact(flags: torch.Tensor) -> torch.Tensor:
    """Ascending indices of the non-zero entries of a flat uint8 CUDA tensor (stable compaction kernel)."""
    n = flags.numel()
    dev = flags.device
    lib = _C.lib()
    scratch = torch.zeros(lib.gsx_compact_scratch_bytes(n), dtype=torch.uint8, device=dev)
    idx = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.gsx_compact_indices(_C.ptr(flags), n, _C.ptr(idx), _C.ptr(cnt), _C.ptr(scratch), 1,
                                     _C.stream_ptr(dev))
    _C.check(rc, "gsx_compact_indices")
    return idx[: int(cnt.item())]  # the table's length is data dependent: this is the API's one host sync


def _check_table(pc2im_bnhw):
    if not torch.is_tensor(pc2im_bnhw):
        raise TypeError("Expected input pc2im_bnhw to be of type torch.Tensor. Got {0} instead.".format(
            type(pc2im_bnhw)))
    if pc2im_bnhw.dtype != torch.int64:
        raise TypeError("Expected input pc2im_bnhw to have dtype of torch.int64 (torch.long), not {0}.".format(
            pc2im_bnhw.dtype))


def _check_table_shape(pc2im_bnhw):
    if pc2im_bnhw.ndim != 2:
        raise ValueError("Expected pc2im_bnhw.ndim of 2. Got {0}.".format(pc2im_bnhw.ndim))
    if pc2im_bnhw.shape[1] != 4:
        raise ValueError("Expected pc2im_bnhw.shape[1] to be 4. Got {0}.".format(pc2im_bnhw.shape[1]))


def _check_pc(pointclouds):
    if not isinstance(pointclouds, Pointclouds):
        raise TypeError("Expected pointclouds to be of type gradslam.Pointclouds. Got {0}.".format(type(pointclouds)))


def _check_batch(pointclouds, rgbdimages):
    if len(rgbdimages) != len(pointclouds):
        raise ValueError("Expected equal batch sizes for pointclouds and rgbdimages. Got {0} and {1} "
                         "respectively.".format(len(pointclouds), len(rgbdimages)))


def find_active_map_points(pointclouds: Pointclouds, rgbdimages: RGBDImages) -> torch.Tensor:
    """int64 (A,4) rows [b, n, h, w] of the map points that project inside the live frame, in (b, n) order
    (fusionutils.py:198-287)."""
    _check_pc(pointclouds)
    _check_frame(rgbdimages)
    device = pointclouds.device
    if not pointclouds.has_points:
        return torch.empty((0, 4), dtype=torch.int64, device=device)
    _check_batch(pointclouds, rgbdimages)
    frames = rgbdimages.to_channels_last()
    B, _, H, W = frames.shape
    st = pointclouds._store
    _C.require_cuda(st["points"], "pointclouds.points")
    width = min(max(pointclouds._bound, 1), pointclouds.capacity)
    flags = torch.empty((B, width), dtype=torch.uint8, device=device)
    hw = torch.empty((B, width), dtype=torch.int32, device=device)
    poses, K = frames.poses.contiguous(), frames.intrinsics.contiguous()
    with torch.cuda.device(device):
        rc = _C.lib().gsx_active_eval(_C.ptr(st["points"]), _C.ptr(pointclouds._counts_dev[pointclouds._cur]),
                                      pointclouds.capacity, width, _C.ptr(poses), 16, _C.ptr(K), 16, B, H, W,
                                      _C.ptr(flags), _C.ptr(hw), _C.stream_ptr(device))
    _C.check(rc, "gsx_active_eval")
    idx = _compact(flags.view(-1))
    pix = hw.view(-1)[idx].to(torch.int64)
    table = torch.stack([idx // width, idx % width, pix // W, pix % W], dim=1)
    if table.shape[0] == 0:
        warnings.warn("No active map points were found")
    return table


def find_similar_map_points(pointclouds: Pointclouds, rgbdimages: RGBDImages, pc2im_bnhw: torch.Tensor,
                            dist_th: Union[float, int], dot_th: Union[float, int]):
    """Rows of the active table whose map point is close to, and has a normal similar to, the frame point of the
    pixel it lands on.  Returns (int64 (S,4), bool (A,)) (fusionutils.py:290-411)."""
    _check_pc(pointclouds)
    if not isinstance(rgbdimages, RGBDImages):
        raise TypeError("Expected rgbdimages to be of type gradslam.RGBDImages. Got {0}.".format(type(rgbdimages)))
    _check_table(pc2im_bnhw)
    if rgbdimages.shape[1] != 1:
        raise ValueError("Expected rgbdimages to have sequence length of 1. Got {0}.".format(rgbdimages.shape[1]))
    _check_table_shape(pc2im_bnhw)
    device = pointclouds.device
    if not pointclouds.has_points or pc2im_bnhw.shape[0] == 0:
        return torch.empty((0, 4), dtype=torch.int64, device=device), torch.empty(0, dtype=torch.bool, device=device)
    _check_batch(pointclouds, rgbdimages)
    if not pointclouds.has_normals:
        raise ValueError("Pointclouds must have normals for finding similar map points, but did not.")
    frames = rgbdimages.to_channels_last()
    B, _, H, W = frames.shape
    table = pc2im_bnhw.contiguous()
    rows = table.shape[0]
    gv, gn = frames.global_vertex_map.contiguous(), frames.global_normal_map.contiguous()
    st = pointclouds._store
    flags = torch.empty(rows, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        rc = _C.lib().gsx_similar_eval(_C.ptr(table), rows, _C.ptr(st["points"]), _C.ptr(st["normals"]),
                                       pointclouds.capacity, _C.ptr(gv), _C.ptr(gn), B, H, W, float(dist_th),
                                       float(dot_th), _C.ptr(flags), _C.stream_ptr(device))
    _C.check(rc, "gsx_similar_eval")
    keep = _compact(flags)
    similar = table[keep]
    if similar.shape[0] == 0:
        warnings.warn("No similar map points were found (despite total {0} active points across the batch)".format(
            rows), RuntimeWarning)
    return similar, flags.bool()


def find_best_unique_correspondences(pointclouds: Pointclouds, rgbdimages: RGBDImages,
                                     pc2im_bnhw: torch.Tensor) -> torch.Tensor:
    """One row per live pixel: among the candidates of a pixel keep the largest confidence count, then the
    smallest ray distance, then the smallest index.  Output sorted by (b, h, w) (fusionutils.py:414-546); the
    reference's torch.unique(dim=0) row sort becomes a per-pixel 128-bit atomic arg-min."""
    _check_pc(pointclouds)
    _check_table(pc2im_bnhw)
    if rgbdimages.shape[1] != 1:
        raise ValueError("Expected rgbdimages to have sequence length of 1. Got {0}.".format(rgbdimages.shape[1]))
    _check_table_shape(pc2im_bnhw)
    device = pointclouds.device
    if not pointclouds.has_points or pc2im_bnhw.shape[0] == 0:
        return torch.empty((0, 4), dtype=torch.int64, device=device)
    _check_batch(pointclouds, rgbdimages)
    if not pointclouds.has_features:
        raise ValueError("Pointclouds must have features for finding best unique correspondences, but did not.")
    frames = rgbdimages.to_channels_last()
    B, _, H, W = frames.shape
    table = pc2im_bnhw.contiguous()
    gv = frames.global_vertex_map.contiguous()
    st = pointclouds._store
    ws = _Workspace.get(device, B, H, W)
    pflags = torch.empty(B * H * W, dtype=torch.uint8, device=device)
    pn = torch.empty(B * H * W, dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        rc = _C.lib().gsx_unique_select(_C.ptr(table), table.shape[0], _C.ptr(st["points"]), _C.ptr(st["features"]),
                                        pointclouds.capacity, _C.ptr(gv), B, H, W, _C.ptr(ws.buf), _C.ptr(pflags),
                                        _C.ptr(pn), _C.stream_ptr(device))
    _C.check(rc, "gsx_unique_select")
    pix = _compact(pflags)
    rem = pix % (H * W)
    return torch.stack([pix // (H * W), pn[pix], rem // W, rem % W], dim=1)


def find_correspondences(pointclouds: Pointclouds, rgbdimages: RGBDImages, dist_th: Union[float, int],
                         dot_th: Union[float, int]) -> torch.Tensor:
    """active -> similar -> best unique (fusionutils.py:549-577)."""
    pc2im_bnhw = find_active_map_points(pointclouds, rgbdimages)
    pc2im_bnhw, _ = find_similar_map_points(pointclouds, rgbdimages, pc2im_bnhw, dist_th, dot_th)
    return find_best_unique_correspondences(pointclouds, rgbdimages, pc2im_bnhw)


def fuse_with_map(pointclouds: Pointclouds, rgbdimages: RGBDImages, pc2im_bnhw: torch.Tensor,
                  sigma: Union[torch.Tensor, float, int], inplace: bool = False) -> Pointclouds:
    """Merges the corresponding points of `pc2im_bnhw` (unique rows) and appends the unmatched valid pixels
    (fusionutils.py:580-722)."""
    _check_pc(pointclouds)
    if not isinstance(rgbdimages, RGBDImages):
        raise TypeError("Expected rgbdimages to be of type gradslam.RGBDImages. Got {0}.".format(type(rgbdimages)))
    _check_table(pc2im_bnhw)
    _check_table_shape(pc2im_bnhw)
    if pointclouds.has_points:
        for what in ("normals", "colors"):
            if not getattr(pointclouds, "has_" + what):
                raise ValueError("Pointclouds must have {} for map fusion, but did not.".format(what))
        if not pointclouds.has_features:
            raise ValueError("Pointclouds must have features (ccounts) for map fusion, but did not.")
    _check_frame(rgbdimages)
    if not inplace:
        pointclouds = pointclouds.clone()
    if _wants_grad(pointclouds, rgbdimages) and rgbdimages.poses is not None:
        return _update_differentiable(pointclouds, rgbdimages, sigma, True, table=pc2im_bnhw)
    frames = rgbdimages.to_channels_last()
    _C.require_cuda(frames.depth_image, "depth_image")
    B, _, H, W = frames.shape
    _prepare_map(pointclouds, frames, True)
    device = pointclouds.device
    if pointclouds._bound > 0 and pc2im_bnhw.shape[0] != 0:
        table = pc2im_bnhw.to(device).contiguous()
        ws = _Workspace.get(device, B, H, W)
        with torch.cuda.device(device):
            rc = _C.lib().gsx_records_from_table(_C.ptr(table), table.shape[0], pointclouds.capacity, B, H, W,
                                                 _C.ptr(ws.buf), _C.stream_ptr(device))
        _C.check(rc, "gsx_records_from_table")
    if frames.poses is not None:
        vmap = nmap = None
    else:
        vmap, nmap = frames.global_vertex_map, frames.global_normal_map
    _launch_merge_append(pointclouds, frames, vmap, nmap, sigma)
    return pointclouds


# --------------------------------------------------------------------------------------------- differentiable mode
def _wants_grad(pointclouds, frames):
    if not torch.is_grad_enabled():
        return False
    ts = [frames.depth_image, frames.poses, frames.rgb_image, frames.intrinsics] + list(pointclouds._store.values())
    return any(torch.is_tensor(t) and t.requires_grad for t in ts)


class _MergeAppendFn(torch.autograd.Function):
    """K4 as one differentiable op: (pre-merge map, frame maps) -> updated map.  forward =
    gsx_fusion_merge_append_fwd on a copy of the map (also records where every pixel went), backward =
    gsx_fusion_merge_append_bwd; both hand-written kernels.  The per-pixel winners must already sit in the fusion
    workspace (K2, or gsx_records_from_table); they are index-only, as in the reference (fusionutils.py:523)."""

    @staticmethod
    def forward(ctx, pack, pts, nrm, col, cc, gv, gn, rgb, vloc):
        pointclouds, frames, sigma = pack
        B, _, H, W = frames.shape
        P = H * W
        dev = pts.device
        bound, cap_in, cap_out = pointclouds._bound, pts.shape[1], pointclouds._bound + P
        outs = []
        for t in (pts, nrm, col, cc):
            if t is None:
                outs.append(None)
                continue
            o = torch.zeros((B, cap_out, t.shape[2]), dtype=torch.float32, device=dev)
            if bound > 0:
                o[:, :bound] = t.detach()[:, :bound]
            outs.append(o)
        gv_c, gn_c, rgb_c, vloc_c = (t.detach().contiguous() for t in (gv, gn, rgb, vloc))
        depth, d_bs = _frame_base(frames.depth_image.detach(), P)
        K = frames.intrinsics.detach().contiguous()
        ws = _Workspace.get(dev, B, H, W)
        counts_in = pointclouds._counts_dev[pointclouds._cur].clone()
        counts_out = pointclouds._counts_dev[pointclouds._cur ^ 1]
        assoc = torch.zeros((B, P), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = _C.lib().gsx_fusion_merge_append_fwd(
                _C.ptr(outs[0]), _C.ptr(outs[1]), _C.ptr(outs[2]), _C.ptr(outs[3]), _C.ptr(counts_in),
                _C.ptr(counts_out), cap_out, _C.ptr(depth), d_bs, _C.ptr(rgb_c), P * 3, _C.ptr(K), 16, _C.ptr(gv_c),
                _C.ptr(gn_c), B, H, W, float(sigma), _C.ptr(ws.buf), ws.next_epochs(1),
                _C.ptr(pointclouds._overflow_flag()), _C.ptr(assoc), _C.stream_ptr(dev))
        _C.check(rc, "gsx_fusion_merge_append_fwd")
        ctx.saved = (assoc, counts_in, pts.detach(), nrm.detach(), col.detach(), None if cc is None else cc.detach(),
                     gv_c, gn_c, rgb_c, vloc_c)
        ctx.dims = (B, H, W, cap_in, cap_out, float(sigma))
        ctx.shapes = (gv.shape, rgb.shape)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_pts, g_nrm, g_col, g_cc):
        assoc, counts_in, pts, nrm, col, cc, gv, gn, rgb, vloc = ctx.saved
        B, H, W, cap_in, cap_out, sigma = ctx.dims
        dev = assoc.device
        gs = [None if g is None else g.contiguous().float() for g in (g_pts, g_nrm, g_col, g_cc)]
        pts_c, nrm_c, col_c = pts.contiguous(), nrm.contiguous(), col.contiguous()
        cc_c = None if cc is None else cc.contiguous()
        d_map = [torch.empty_like(t) for t in (pts_c, nrm_c, col_c)] + [None if cc_c is None else torch.empty_like(cc_c)]
        d_frame = [torch.empty((B, 1, H, W, 3), dtype=torch.float32, device=dev) for _ in range(4)]
        with torch.cuda.device(dev):
            rc = _C.lib().gsx_fusion_merge_append_bwd(
                _C.ptr(assoc), _C.ptr(counts_in), _C.ptr(pts_c), _C.ptr(nrm_c), _C.ptr(col_c), _C.ptr(cc_c), cap_in,
                _C.ptr(gs[0]), _C.ptr(gs[1]), _C.ptr(gs[2]), _C.ptr(gs[3]), cap_out, _C.ptr(gv), _C.ptr(gn),
                _C.ptr(rgb), _C.ptr(vloc), B, H, W, sigma, _C.ptr(d_map[0]), _C.ptr(d_map[1]), _C.ptr(d_map[2]),
                _C.ptr(d_map[3]), _C.ptr(d_frame[0]), _C.ptr(d_frame[1]), _C.ptr(d_frame[2]), _C.ptr(d_frame[3]),
                _C.stream_ptr(dev))
        _C.check(rc, "gsx_fusion_merge_append_bwd")
        gv_shape, rgb_shape = ctx.shapes
        return (None, d_map[0], d_map[1], d_map[2], d_map[3], d_frame[0].view(gv_shape), d_frame[1].view(gv_shape),
                d_frame[2].view(rgb_shape), d_frame[3].view(gv_shape))


def _update_differentiable(pointclouds, frames, sigma, with_features, dist_th=None, dot_th=None, table=None):
    """Map update when a gradient is requested: K1 (differentiable op) -> association (K2 kernel, or the rows of
    `table`; index-only) -> K4 (differentiable op), out of place so the pre-merge map survives for the backward.
    Values equal the in-place kernel path bit for bit."""
    frames = frames.to_channels_last()
    _C.require_cuda(frames.depth_image, "depth_image")
    B, _, H, W = frames.shape
    P = H * W
    if not pointclouds.has_points:
        pointclouds.device = frames.device
        pointclouds._allocate(B, 1, 1 if with_features else 0)
    elif len(pointclouds) != B:
        raise ValueError("Expected equal batch sizes for pointclouds and rgbdimages. Got {0} and {1} "
                         "respectively.".format(len(pointclouds), B))
    dev = pointclouds.device
    st = pointclouds._store
    gv, gn, vloc = frames.global_vertex_map, frames.global_normal_map, frames.vertex_map  # K1, carries its backward
    ws = _Workspace.get(dev, B, H, W)
    if pointclouds._bound > 0 and table is not None and table.shape[0] != 0:
        table = table.to(dev).contiguous()
        with torch.cuda.device(dev):
            rc = _C.lib().gsx_records_from_table(_C.ptr(table), table.shape[0], pointclouds.capacity, B, H, W,
                                                 _C.ptr(ws.buf), _C.stream_ptr(dev))
        _C.check(rc, "gsx_records_from_table")
    elif pointclouds._bound > 0 and dist_th is not None:
        pts, nrm, ccs = (st[k].detach().contiguous() for k in ("points", "normals", "features"))
        K = frames.intrinsics.detach().contiguous()
        poses = frames.poses.detach().contiguous()
        depth, d_bs = _frame_base(frames.depth_image.detach(), P)
        gvd, gnd = gv.detach().contiguous(), gn.detach().contiguous()
        with torch.cuda.device(dev):
            rc = _C.lib().gsx_fusion_project_select(
                _C.ptr(pts), _C.ptr(nrm), _C.ptr(ccs), _C.ptr(pointclouds._counts_dev[pointclouds._cur]),
                pointclouds.capacity, pointclouds._bound, _C.ptr(poses), 16, _C.ptr(K), 16, _C.ptr(depth), d_bs,
                _C.ptr(gvd), _C.ptr(gnd), B, H, W, float(dist_th), float(dot_th), _C.ptr(ws.buf), _C.stream_ptr(dev))
        _C.check(rc, "gsx_fusion_project_select")
    outs = _MergeAppendFn.apply((pointclouds, frames, sigma), st["points"], st["normals"], st["colors"],
                                st["features"], gv, gn, frames.rgb_image, vloc)
    for key, o in zip(("points", "normals", "colors", "features"), outs):
        st[key] = o
    pointclouds._uninit = False
    pointclouds._mark_device_updated(pointclouds._bound + P)
    return pointclouds


# --------------------------------------------------------------------------------------------- public ops
def update_map_aggregate(pointclouds: Pointclouds, rgbdimages: RGBDImages, inplace: bool = False) -> Pointclouds:
    """Appends every valid live-frame pixel to the maps (fusionutils.py:725-758)."""
    if not isinstance(pointclouds, Pointclouds):
        raise TypeError("Expected pointclouds to be of type gradslam.Pointclouds. Got {0}.".format(type(pointclouds)))
    if not isinstance(rgbdimages, RGBDImages):
        raise TypeError("Expected rgbdimages to be of type gradslam.RGBDImages. Got {0}.".format(type(rgbdimages)))
    if not inplace:
        pointclouds = pointclouds.clone()
    if _wants_grad(pointclouds, rgbdimages):
        _check_frame(rgbdimages)
        return _update_differentiable(pointclouds, rgbdimages, 0.6, pointclouds.has_features)
    return _append_valid_pixels(pointclouds, rgbdimages, True)


def update_map_fusion(pointclouds: Pointclouds, rgbdimages: RGBDImages, dist_th: Union[float, int],
                      dot_th: Union[float, int], sigma: Union[torch.Tensor, float, int],
                      inplace: bool = False) -> Pointclouds:
    """PointFusion update of the maps with one live frame (fusionutils.py:761-789)."""
    if not isinstance(pointclouds, Pointclouds):
        raise TypeError("Expected pointclouds to be of type gradslam.Pointclouds. Got {0}.".format(type(pointclouds)))
    _check_frame(rgbdimages)
    if not inplace:
        pointclouds = pointclouds.clone()
    if _wants_grad(pointclouds, rgbdimages):
        if rgbdimages.poses is None:
            raise ValueError("rgbdimages must have poses for map fusion")
        if pointclouds.has_points:
            for what in ("normals", "colors", "features"):
                if not getattr(pointclouds, "has_" + what):
                    raise ValueError("Pointclouds must have {} for map fusion, but did not.".format(what))
        return _update_differentiable(pointclouds, rgbdimages, sigma, True, dist_th, dot_th)
    return _fused_update(pointclouds, rgbdimages, dist_th, dot_th, sigma)
