"""CPU oracle for the PointFusion / ICPSLAM hot path (TEST INFRASTRUCTURE — NOT PRODUCT CODE).

This file is a from-scratch CPU restatement (torch-CPU tensor ops + one small C
routine for the exact 1-NN) of the algorithm that gradslam/gradslam runs for

    RGBDImages vertex/normal maps      gradslam/structures/rgbdimages.py:643-762
    projective data association        gradslam/slam/fusionutils.py:198-546
    confidence-weighted surfel fusion  gradslam/slam/fusionutils.py:16-73, 580-722
    point-to-plane ICP / gradICP       gradslam/odometry/icputils.py:22-545
    downsampling for ICP               gradslam/odometry/icputils.py:548-669
    SE(3) exponential                  gradslam/geometry/se3utils.py:11-115
    sequence drivers                   gradslam/slam/icpslam.py:99-264, slam/pointfusion.py:107-112

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it, and only as the checker / timed
CPU baseline.  The product package `gradslam_b200` never imports it; the
product has no CPU fallback.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the UNMODIFIED
reference (imported from /root/reference under the shims in
tests/golden/ref_loader.py) and freezes its outputs; `tests/test_oracle_golden.py`
checks this oracle against those fixtures and against the reference's own
golden vectors (tests/data/msrd_b2s3) and known-answer tests.  Backward passes
are pinned too: `tests/golden/make_golden_grad.py` records the reference's
autograd gradients (tests/golden/ref_grad.npz) and the same test file compares
this oracle's autograd with them.

Canonical arithmetic.  The reference computes 3-term dot products through
einsum/bmm, whose rounding order is whatever the BLAS picks.  So that the CUDA
kernels can be BIT-exact on every decision (thresholds, rounding to pixels,
argmin keys) the oracle fixes one order:  every product and sum is rounded
separately to float32 (no FMA) and 3-term sums associate left to right,
`(a*x + b*y) + c*z`, then `+ t`; square roots and the confidence weight's exp are
taken in float64 and rounded once (correctly rounded float32 results).  The one
place where the reference's own rounding is known and matters for DECISIONS is the
normal estimate: its CPU torch.cross evaluates `a*b - c*d` as fma(a, b, -(c*d)) and
Tensor.norm as sqrt(fma(z,z,fma(y,y,x*x))) - at pixels whose right and lower
neighbours are both missing the two differences are equal and the contracted cross
product is rounding residue, not 0, which after normalisation decides whether a map
point matches.  The oracle (oracle/normal_fma.c) and the kernels (gsx_common.cuh)
use exactly that arithmetic, so local normal maps are bit-identical to the
reference's.  Remaining differences from the reference are at the 1-ulp level and
are covered by the tolerances in the golden tests.

One deliberate deviation: the reference's merge rewrites EVERY map point as
`(c*p) * (1/c)` each frame (fusionutils.py:682-699 operates on the whole padded
map), which perturbs unmatched points by rounding noise.  Mathematically this
is the identity; the oracle and the CUDA path leave unmatched points untouched.

The third-party 1-NN (`chamferdist.chamfer.knn_points`, pinned
chamferdist==1.0.0, call site odometry/icputils.py:200; source not vendored) is
restated in oracle/knn1.c as the published brute-force algorithm: for every
query scan all targets, squared L2, keep the first minimum.
"""
import ctypes
import math
import os
import subprocess
from collections import namedtuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
F32 = torch.float32


# ------------------------------------------------------------------------------------------
# small rigid-transform helpers (kornia.geometry.linalg restatements; call sites
# slam/fusionutils.py:249, slam/icpslam.py:245)
# ------------------------------------------------------------------------------------------
def _sqrt32(x):
    """Correctly rounded float32 square root.  torch's CPU float32 sqrt is NOT always correctly rounded
    (observed 1-ulp misses with torch 2.11 / AVX2); sqrt in float64 rounded once to float32 is exact and is
    what CUDA's sqrtf (-prec-sqrt=true) and numpy return."""
    return torch.sqrt(x.double()).float()


def _dot3(a0, a1, a2, b0, b1, b2):
    """(a0*b0 + a1*b1) + a2*b2 with every op rounded separately."""
    return (a0 * b0 + a1 * b1) + a2 * b2


def rigid_inverse(T):
    """[R^T, (-R^T) t] with bottom row [0,0,0,1].  T: (...,4,4)."""
    R = T[..., :3, :3]
    t = T[..., :3, 3]
    out = torch.zeros_like(T)
    Rt = R.transpose(-1, -2)
    out[..., :3, :3] = Rt
    nRt = -Rt
    for i in range(3):
        out[..., i, 3] = _dot3(nRt[..., i, 0], nRt[..., i, 1], nRt[..., i, 2], t[..., 0], t[..., 1], t[..., 2])
    out[..., 3, 3] = 1.0
    return out


def rigid_compose(A, B):
    """A·B for rigid 4x4 (rotation/translation blocks only, bottom row forced)."""
    out = torch.zeros_like(A)
    for i in range(3):
        for j in range(3):
            out[..., i, j] = _dot3(A[..., i, 0], A[..., i, 1], A[..., i, 2], B[..., 0, j], B[..., 1, j], B[..., 2, j])
        out[..., i, 3] = _dot3(A[..., i, 0], A[..., i, 1], A[..., i, 2], B[..., 0, 3], B[..., 1, 3], B[..., 2, 3]) + A[..., i, 3]
    out[..., 3, 3] = 1.0
    return out


def mat4_mul(A, B):
    """Plain 4x4 product, left-to-right accumulation (torch.mm on 4x4 in icputils.py:362,543)."""
    out = torch.zeros_like(A)
    for i in range(4):
        for j in range(4):
            acc = A[..., i, 0] * B[..., 0, j]
            for k in range(1, 4):
                acc = acc + A[..., i, k] * B[..., k, j]
            out[..., i, j] = acc
    return out


def rigid_apply(T, P):
    """R·p + t for P (...,N,3) and T (...,4,4) broadcast over N."""
    x, y, z = P[..., 0], P[..., 1], P[..., 2]
    cols = []
    for i in range(3):
        r0 = T[..., i, 0].unsqueeze(-1)
        r1 = T[..., i, 1].unsqueeze(-1)
        r2 = T[..., i, 2].unsqueeze(-1)
        t = T[..., i, 3].unsqueeze(-1)
        cols.append(_dot3(r0, r1, r2, x, y, z) + t)
    return torch.stack(cols, -1)


# ------------------------------------------------------------------------------------------
# K1: depth -> vertex / normal maps   (structures/rgbdimages.py:643-762, geometry/projutils.py:405-450)
# ------------------------------------------------------------------------------------------
def inverse_intrinsics(K, eps=1e-6):
    Kinv = torch.zeros_like(K)
    fx, fy, cx, cy = K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2]
    Kinv[..., 0, 0] = 1.0 / (fx + eps)
    Kinv[..., 1, 1] = 1.0 / (fy + eps)
    Kinv[..., 0, 2] = -1.0 * cx / (fx + eps)
    Kinv[..., 1, 2] = -1.0 * cy / (fy + eps)
    Kinv[..., 2, 2] = 1
    Kinv[..., -1, -1] = 1
    return Kinv


def frame_maps(depth, K, poses=None):
    """depth (B,L,H,W,1), K (B,1,4,4), poses (B,L,4,4)|None -> dict of (B,L,H,W,3) maps + valid mask."""
    B, L, H, W, _ = depth.shape
    Kinv = inverse_intrinsics(K)  # (B,1,4,4)
    u = torch.arange(W, dtype=F32).view(1, 1, 1, W)
    v = torch.arange(H, dtype=F32).view(1, 1, H, 1)
    d = depth[..., 0]
    valid = d > 0
    vf = valid.to(F32)
    k00 = Kinv[..., 0, 0].view(B, 1, 1, 1)
    k02 = Kinv[..., 0, 2].view(B, 1, 1, 1)
    k11 = Kinv[..., 1, 1].view(B, 1, 1, 1)
    k12 = Kinv[..., 1, 2].view(B, 1, 1, 1)
    vx = ((k00 * u + k02) * d) * vf
    vy = ((k11 * v + k12) * d) * vf
    vz = d * vf
    vert = torch.stack([vx.expand(B, L, H, W), vy.expand(B, L, H, W), vz], -1).contiguous()

    # forward differences; last column / row copies its neighbour (rgbdimages.py:724-731)
    dh = torch.zeros_like(vert)
    dv = torch.zeros_like(vert)
    dh[..., :, :-1, :] = vert[..., :, 1:, :] - vert[..., :, :-1, :]
    dv[..., :-1, :, :] = vert[..., 1:, :, :] - vert[..., :-1, :, :]
    dh[..., :, -1, :] = dh[..., :, -2, :]
    dv[..., -1, :, :] = dv[..., -2, :, :]
    # torch.cross / Tensor.norm as the reference's CPU build rounds them (oracle/normal_fma.c)
    cross, nrm = cross_norm_fma(dh, dv)
    cx, cy, cz = cross[..., 0], cross[..., 1], cross[..., 2]
    den = torch.where(nrm == 0, torch.ones_like(nrm), nrm)
    normal = torch.stack([(cx / den) * vf, (cy / den) * vf, (cz / den) * vf], -1)

    if poses is None:
        gvert, gnormal = vert.clone(), normal.clone()
    else:
        R = poses[..., :3, :3]
        t = poses[..., :3, 3]
        gv, gn = [], []
        for i in range(3):
            r0 = R[..., i, 0].view(B, L, 1, 1)
            r1 = R[..., i, 1].view(B, L, 1, 1)
            r2 = R[..., i, 2].view(B, L, 1, 1)
            ti = t[..., i].view(B, L, 1, 1)
            gv.append((_dot3(r0, r1, r2, vert[..., 0], vert[..., 1], vert[..., 2]) + ti) * vf)
            gn.append(_dot3(r0, r1, r2, normal[..., 0], normal[..., 1], normal[..., 2]))
        gvert, gnormal = torch.stack(gv, -1), torch.stack(gn, -1)
    return {"vertex": vert, "normal": normal, "gvertex": gvert, "gnormal": gnormal, "valid": valid}


# ------------------------------------------------------------------------------------------
# map container (list form, like gradslam.Pointclouds after append_points; pointclouds.py:1117-1237)
# ------------------------------------------------------------------------------------------
class SurfelMap:
    """Per-element lists of (N_b,3) points / normals / colors and (N_b,1) confidence counts."""

    def __init__(self, points=None, normals=None, colors=None, ccounts=None):
        self.points = points
        self.normals = normals
        self.colors = colors
        self.ccounts = ccounts

    @property
    def has_points(self):
        return self.points is not None

    @property
    def B(self):
        return 0 if self.points is None else len(self.points)

    def counts(self):
        return [int(p.shape[0]) for p in self.points]

    def padded(self):
        """list -> zero-padded (B,N,C) tensors (structutils.py:47-86)."""
        def pad(lst):
            if lst is None:
                return None
            n = max(x.shape[0] for x in lst)
            out = torch.zeros(len(lst), n, lst[0].shape[1], dtype=F32)
            for i, x in enumerate(lst):
                out[i, : x.shape[0]] = x
            return out
        return pad(self.points), pad(self.normals), pad(self.colors), pad(self.ccounts)

    def nonpad_mask(self):
        c = self.counts()
        n = max(c)
        return torch.arange(n).view(1, -1) < torch.tensor(c).view(-1, 1)

    def clone(self):
        if not self.has_points:
            return SurfelMap()
        cl = lambda l: None if l is None else [x.clone() for x in l]
        return SurfelMap(cl(self.points), cl(self.normals), cl(self.colors), cl(self.ccounts))

    def append(self, other):
        if not other.has_points:
            return self
        if not self.has_points:
            o = other.clone()
            self.points, self.normals, self.colors, self.ccounts = o.points, o.normals, o.colors, o.ccounts
            return self
        cat = lambda a, b: None if a is None else [torch.cat([x, y], 0) for x, y in zip(a, b)]
        self.points = cat(self.points, other.points)
        self.normals = cat(self.normals, other.normals)
        self.colors = cat(self.colors, other.colors)
        self.ccounts = cat(self.ccounts, other.ccounts)
        return self


# ------------------------------------------------------------------------------------------
# K2: active / similar map points   (slam/fusionutils.py:198-411)
# ------------------------------------------------------------------------------------------
def project_map(points_padded, pose, K):
    """World points -> (u, v, z_cam) in the live camera.  pose (B,4,4) camera-to-world, K (B,4,4).
    transform: pointclouds.py:526-573; projection: projutils.py:92-238 (z==0 divides by 1)."""
    Tinv = rigid_inverse(pose)
    q = rigid_apply(Tinv, points_padded)  # (B,N,3)
    x, y, z = q[..., 0], q[..., 1], q[..., 2]
    row = lambda i: (((K[:, i, 0:1] * x + K[:, i, 1:2] * y) + K[:, i, 2:3] * z) + K[:, i, 3:4])
    px, py, pz = row(0), row(1), row(2)
    den = torch.where(pz != 0, pz, torch.ones_like(pz))
    return px / den, py / den, z


def find_active_map_points(smap, pose, K, H, W):
    """-> int64 (A,4) rows [b, n, h, w] in (b, n) order."""
    if not smap.has_points:
        return torch.empty((0, 4), dtype=torch.int64)
    pts = smap.padded()[0]
    u, v, z = project_map(pts, pose, K)
    lo = torch.tensor(-1e-3, dtype=F32)
    inframe = (
        (u > lo) & (u < torch.tensor(W - 0.999, dtype=F32))
        & (v > lo) & (v < torch.tensor(H - 0.999, dtype=F32))
        & (z > 0) & smap.nonpad_mask()
    )
    w = u.round().long().clamp(0, W - 1)
    h = v.round().long().clamp(0, H - 1)
    B, N = u.shape
    bb = torch.arange(B).view(B, 1).expand(B, N)
    nn = torch.arange(N).view(1, N).expand(B, N)
    table = torch.stack([bb, nn, h, w], -1)
    return table[inframe]


def find_similar_map_points(smap, gvertex, gnormal, pc2im, dist_th, dot_th):
    """gvertex/gnormal (B,H,W,3).  -> (S,4) rows kept, bool (A,) mask."""
    if not smap.has_points or pc2im.shape[0] == 0:
        return torch.empty((0, 4), dtype=torch.int64), torch.empty(0, dtype=torch.bool)
    pts, nrm, _, _ = smap.padded()
    b, n, h, w = pc2im.unbind(1)
    fp = gvertex[b, h, w]
    fn = gnormal[b, h, w]
    mp = pts[b, n]
    mn = nrm[b, n]
    d = fp - mp
    dist = _sqrt32((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    dot = (fn[:, 0] * mn[:, 0] + fn[:, 1] * mn[:, 1]) + fn[:, 2] * mn[:, 2]
    mask = (dist < torch.tensor(dist_th, dtype=F32)) & (dot > torch.tensor(dot_th, dtype=F32))
    return pc2im[mask], mask


def unique_sort_keys(smap, gvertex, pc2im):
    """float32 (S,6) rows [b, h, w, 1/(cc+1e-20), ray_dist, n]  (fusionutils.py:491-517)."""
    pts, _, _, cc = smap.padded()
    b, n, h, w = pc2im.unbind(1)
    inv_cc = 1 / (cc[b, n] + 1e-20)  # (S,1)
    d = pts[b, n] - gvertex[b, h, w]
    ray = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).unsqueeze(1)
    return torch.cat([b.unsqueeze(1).float(), h.unsqueeze(1).float(), w.unsqueeze(1).float(),
                      inv_cc, ray, n.unsqueeze(1).float()], -1)


def find_best_unique_correspondences(smap, gvertex, pc2im):
    """Per (b,h,w): largest ccount, then smallest ray distance, then smallest n.  Output sorted by (b,h,w).
    Uses torch.unique(dim=0) exactly like the reference (fusionutils.py:522) — it is the
    reference's dominant CPU cost and is kept so the CPU baseline is faithful."""
    if not smap.has_points or pc2im.shape[0] == 0:
        return torch.empty((0, 4), dtype=torch.int64)
    crit = unique_sort_keys(smap, gvertex, pc2im)
    srt = torch.unique(crit, dim=0)
    first = torch.ones(srt.shape[0], dtype=torch.bool)
    first[1:] = (srt[1:, :3] != srt[:-1, :3]).any(-1)
    keep = srt[first]
    return torch.stack([keep[:, 0].long(), keep[:, 5].long(), keep[:, 1].long(), keep[:, 2].long()], -1)


def find_correspondences(smap, maps, pose, K, dist_th, dot_th):
    B, H, W = maps["gvertex"].shape[0], maps["gvertex"].shape[-3], maps["gvertex"].shape[-2]
    gv, gn = maps["gvertex"][:, 0], maps["gnormal"][:, 0]
    t = find_active_map_points(smap, pose, K, H, W)
    t, _ = find_similar_map_points(smap, gv, gn, t, dist_th, dot_th)
    return find_best_unique_correspondences(smap, gv, t)


# ------------------------------------------------------------------------------------------
# K4: merge matched + append new   (slam/fusionutils.py:16-73, 580-722)
# ------------------------------------------------------------------------------------------
def get_alpha(vertex, sigma, eps=1e-7):
    """clamp(exp(-(x^2+y^2+z^2) / (2 sigma^2)), eps, 1.01) on the LOCAL vertex; (...,3)->(...,1)."""
    s = (vertex[..., 0] * vertex[..., 0] + vertex[..., 1] * vertex[..., 1]) + vertex[..., 2] * vertex[..., 2]
    # exp in float64, rounded once to float32 (the correctly rounded result; see confidence_alpha in gsx_fusion.cu)
    a = torch.exp((-s / torch.tensor(2 * (sigma ** 2), dtype=F32)).double()).float()
    return torch.clamp(a, min=eps, max=1.01).unsqueeze(-1)


def fuse_with_map(smap, maps, rgb, pc2im, sigma):
    """maps from frame_maps() with L==1; rgb (B,1,H,W,3); pc2im (U,4) unique.  Returns a NEW SurfelMap."""
    gv, gn, col = maps["gvertex"][:, 0], maps["gnormal"][:, 0], rgb[:, 0]
    alpha = get_alpha(maps["vertex"][:, 0], sigma)  # (B,H,W,1)
    valid = maps["valid"][:, 0]
    B = gv.shape[0]
    out = smap.clone()
    new_mask = torch.ones_like(valid)
    if smap.has_points and pc2im.shape[0] != 0:
        b, n, h, w = pc2im.unbind(1)
        for i in range(B):
            sel = b == i
            ni, hi, wi = n[sel], h[sel], w[sel]
            cc = out.ccounts[i][ni]
            a = alpha[i, hi, wi]
            tot = cc + a
            inv = 1 / torch.where(tot == 0, torch.ones_like(tot), tot)
            out.points[i][ni] = ((cc * out.points[i][ni]) + (a * gv[i, hi, wi])) * inv
            out.normals[i][ni] = ((cc * out.normals[i][ni]) + (a * gn[i, hi, wi])) * inv
            out.colors[i][ni] = ((cc * out.colors[i][ni]) + (a * col[i, hi, wi])) * inv
            out.ccounts[i][ni] = tot
        new_mask[b, h, w] = False
    new_mask = new_mask & valid
    fresh = SurfelMap(
        [gv[i][new_mask[i]] for i in range(B)],
        [gn[i][new_mask[i]] for i in range(B)],
        [col[i][new_mask[i]] for i in range(B)],
        [alpha[i][new_mask[i]] for i in range(B)],
    )
    return out.append(fresh)


def update_map_fusion(smap, maps, rgb, pose, K, dist_th, dot_th, sigma):
    t = find_correspondences(smap, maps, pose, K, dist_th, dot_th)
    return fuse_with_map(smap, maps, rgb, t, sigma)


def update_map_aggregate(smap, maps, rgb):
    """Append every valid pixel (slam/fusionutils.py:725-758, structures/utils.py:7-57); no ccounts."""
    gv, gn, col, valid = maps["gvertex"][:, 0], maps["gnormal"][:, 0], rgb[:, 0], maps["valid"][:, 0]
    B = gv.shape[0]
    fresh = SurfelMap([gv[i][valid[i]] for i in range(B)], [gn[i][valid[i]] for i in range(B)],
                      [col[i][valid[i]] for i in range(B)], None)
    return smap.clone().append(fresh)


# ------------------------------------------------------------------------------------------
# K5: exact 1-NN (C)   — chamferdist.chamfer.knn_points restatement, see oracle/knn1.c
# ------------------------------------------------------------------------------------------
_knn_lib = None


def build_c(force=False):
    """Compile oracle/*.c -> oracle/_build/libgsx_oracle.so (gcc, no FMA contraction)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libgsx_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("knn1.c", "normal_fma.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp"] + srcs + ["-o", so, "-lm"]
        subprocess.run(cmd, check=True)
    return so


def _lib():
    global _knn_lib
    if _knn_lib is None:
        _knn_lib = ctypes.CDLL(build_c())
        _knn_lib.gsx_oracle_knn1.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _knn_lib.gsx_oracle_knn1.restype = None
        _knn_lib.gsx_oracle_cross_norm_fma.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                        ctypes.c_void_p, ctypes.c_void_p]
        _knn_lib.gsx_oracle_cross_norm_fma.restype = None
    return _knn_lib


class _CrossNormFma(torch.autograd.Function):
    """(a x b, |a x b|) with the forward rounded by oracle/normal_fma.c and the analytic backward
    (d(a x b) = da x b + a x db;  d|c| = c . dc / |c|), so the oracle stays differentiable for the gradient tests."""

    @staticmethod
    def forward(ctx, a, b):
        shape = a.shape
        a2 = a.detach().contiguous().float().view(-1, 3)
        b2 = b.detach().contiguous().float().view(-1, 3)
        c = torch.empty_like(a2)
        n = torch.empty(a2.shape[0], dtype=F32)
        _lib().gsx_oracle_cross_norm_fma(a2.data_ptr(), b2.data_ptr(), a2.shape[0], c.data_ptr(), n.data_ptr())
        c, n = c.view(shape), n.view(shape[:-1])
        ctx.save_for_backward(a, b, c, n)
        return c, n

    @staticmethod
    def backward(ctx, g_c, g_n):
        a, b, c, n = ctx.saved_tensors
        den = torch.where(n == 0, torch.ones_like(n), n).unsqueeze(-1)
        g = g_c + g_n.unsqueeze(-1) * c / den * (n != 0).unsqueeze(-1).to(c.dtype)
        return torch.cross(b, g, dim=-1), torch.cross(g, a, dim=-1)


def cross_norm_fma(a, b):
    """a x b and its length for (...,3) float32 tensors, rounded like the reference's torch.cross / Tensor.norm on
    the CPU (oracle/normal_fma.c; rgbdimages.py:733-734)."""
    return _CrossNormFma.apply(a, b)


def knn1(src, tgt, threads=0):
    """src (Ns,3), tgt (Nt,3) float32 -> (squared dist (Ns,), idx int64 (Ns,)); ties -> lowest index."""
    src = src.contiguous().float()
    tgt = tgt.contiguous().float()
    d = torch.empty(src.shape[0], dtype=F32)
    i = torch.empty(src.shape[0], dtype=torch.int64)
    _lib().gsx_oracle_knn1(src.data_ptr(), src.shape[0], tgt.data_ptr(), tgt.shape[0], d.data_ptr(), i.data_ptr(),
                           int(threads))
    return d, i


# ------------------------------------------------------------------------------------------
# K6/K7: point-to-plane ICP   (odometry/icputils.py:22-545, geometry/se3utils.py)
# ------------------------------------------------------------------------------------------
def gauss_newton_solve(src, tgt, tgt_normals, dist_thresh=None):
    """src (Ns,3), tgt (Nt,3) -> A (Nsf,6), b (Nsf,1), idx (Nsf,).  Note the reference compares the
    SQUARED nn distance with the un-squared threshold (icputils.py:206); reproduced."""
    d2, idx = knn1(src, tgt)
    keep = torch.ones_like(d2, dtype=torch.bool) if dist_thresh is None else d2 < torch.tensor(dist_thresh, dtype=F32)
    idx = idx[keep]
    s = src[keep]
    sx, sy, sz = s[:, 0:1], s[:, 1:2], s[:, 2:3]
    p = tgt[idx]
    n = tgt_normals[idx]
    dx, dy, dz = p[:, 0:1], p[:, 1:2], p[:, 2:3]
    nx, ny, nz = n[:, 0:1], n[:, 1:2], n[:, 2:3]
    A = torch.cat([nx, ny, nz, nz * sy - ny * sz, nx * sz - nz * sx, ny * sx - nx * sy], 1)
    b = nx * (dx - sx) + ny * (dy - sy) + nz * (dz - sz)
    return A, b, idx


def solve_linear_system(A, b, damp=1e-8):
    damp = damp if torch.is_tensor(damp) else torch.tensor(damp, dtype=A.dtype)
    At = A.t()
    AtA = At @ A + torch.eye(A.shape[1]) * damp
    return torch.inverse(AtA) @ (At @ b)


def so3_hat(w):
    z = torch.zeros((), dtype=w.dtype)
    return torch.stack([torch.stack([z, -w[2], w[1]]), torch.stack([w[2], z, -w[0]]), torch.stack([-w[1], w[0], z])])


def se3_exp(xi):
    """xi (6,) or (6,1) = (v, omega) -> 4x4.  Small-angle branch uses I + hat for BOTH R and V (se3utils.py:91-93)."""
    xi = xi.reshape(6)
    v, w = xi[:3], xi[3:]
    W = so3_hat(w)
    theta = _sqrt32((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2])
    I = torch.eye(3, dtype=xi.dtype)
    if theta < 1e-6:
        R = I + W
        V = I + W
    else:
        s, c = theta.sin(), theta.cos()
        W2 = W @ W
        Ac = s / theta
        Bc = (1 - c) / (theta * theta)
        Cc = (theta - s) / (theta * theta * theta)
        R = I + Ac * W + Bc * W2
        V = I + Bc * W + Cc * W2
    t = V @ v.view(3, 1)
    T = torch.eye(4, dtype=xi.dtype)
    T[:3, :3] = R
    T[:3, 3:] = t
    return T


def _apply(T, P):
    return rigid_apply(T, P)


def point_to_plane_icp(src, tgt, tgt_normals, T0, numiters=20, damp=1e-8, dist_thresh=None):
    """LM with accept/reject (icputils.py:235-367).  src (Ns,3) -> (T (4,4), last idx)."""
    damp = torch.tensor(damp, dtype=F32)
    src = _apply(T0, src)
    T = T0
    idx = None
    for _ in range(numiters):
        A, b, idx = gauss_newton_solve(src, tgt, tgt_normals, dist_thresh)
        xi = solve_linear_system(A, b, damp)
        dT = se3_exp(xi)
        err = torch.dot(b[:, 0], b[:, 0])
        one = _apply(dT, src)
        _, b1, _ = gauss_newton_solve(one, tgt, tgt_normals, dist_thresh)
        new_err = torch.dot(b1[:, 0], b1[:, 0])
        if new_err < err:
            src = one
            damp = damp / 2
            T = mat4_mul(dT, T)
        else:
            damp = damp * 2
    return T, idx


def point_to_plane_gradicp(src, tgt, tgt_normals, T0, numiters=20, damp=1e-8, dist_thresh=None,
                           lambda_max=2.0, B=1.0, B2=1.0, nu=200.0):
    """gradLM (icputils.py:370-545): smooth damping / step gates, clamp(+-70)."""
    damp = torch.tensor(damp, dtype=F32)
    lambda_min = 1 / lambda_max
    src = _apply(T0, src)
    T = T0
    idx = None
    for _ in range(numiters):
        A, b, idx = gauss_newton_solve(src, tgt, tgt_normals, dist_thresh)
        xi = solve_linear_system(A, b, damp)
        dT = se3_exp(xi)
        err = torch.dot(b[:, 0], b[:, 0])
        one = _apply(dT, src)
        _, b1, _ = gauss_newton_solve(one, tgt, tgt_normals, dist_thresh)
        new_err = torch.dot(b1[:, 0], b1[:, 0])
        diff = (new_err - err).clamp(-70.0, 70.0)
        damp = damp * (lambda_min + (lambda_max - lambda_min) / (1 + torch.exp(-B * diff)))
        sig = 1 / ((1 + torch.exp(-B2 * diff)) ** (1 / nu))
        dT = se3_exp(sig * xi)
        src = _apply(dT, src)
        T = mat4_mul(dT, T)
    return T, idx


def downsample_frame(maps, ds):
    """Strided subsample of GLOBAL maps + valid mask (icputils.py:623-669) -> per-b lists (points, normals)."""
    gv, gn, valid = maps["gvertex"][:, 0], maps["gnormal"][:, 0], maps["valid"][:, 0]
    m = valid[:, ::ds, ::ds]
    B = gv.shape[0]
    return ([gv[b, ::ds, ::ds][m[b]] for b in range(B)], [gn[b, ::ds, ::ds][m[b]] for b in range(B)])


def downsample_map(smap, pc2im, ds):
    """Keep active rows whose pixel lies on the ds lattice (icputils.py:548-620) -> per-b lists."""
    t = pc2im[pc2im[:, 2] % ds == 0]
    t = t[t[:, 3] % ds == 0]
    B = smap.B
    pts = [smap.points[b][t[t[:, 0] == b][:, 1]] for b in range(B)]
    nrm = [smap.normals[b][t[t[:, 0] == b][:, 1]] for b in range(B)]
    return pts, nrm


def odometry(smap, live_maps_at_prev_pose, prev_pose, K, H, W, odom, ds, icp_kwargs):
    """ICPSLAM._localize for odom in {icp, gradicp} (slam/icpslam.py:238-247) -> new pose (B,4,4)."""
    f_pts, _ = downsample_frame(live_maps_at_prev_pose, ds)
    table = find_active_map_points(smap, prev_pose, K, H, W)
    m_pts, m_nrm = downsample_map(smap, table, ds)
    fn = point_to_plane_icp if odom == "icp" else point_to_plane_gradicp
    Ts = []
    for b in range(smap.B):
        T, _ = fn(f_pts[b], m_pts[b], m_nrm[b], torch.eye(4), **icp_kwargs)
        Ts.append(T)
    return rigid_compose(torch.stack(Ts), prev_pose)


# ------------------------------------------------------------------------------------------
# drivers   (slam/icpslam.py:99-178, slam/pointfusion.py)
# ------------------------------------------------------------------------------------------
SlamResult = namedtuple("SlamResult", "map poses")


def run_slam(rgb, depth, K, poses=None, *, mode="pointfusion", odom="gt", dist_th=0.05, angle_th=20.0, sigma=0.6,
             dsratio=4, numiters=20, damp=1e-8, dist_thresh=None, lambda_max=2.0, B=1.0, B2=1.0, nu=200.0):
    """rgb (B,L,H,W,3), depth (B,L,H,W,1), K (B,1,4,4), poses (B,L,4,4)|None.
    mode 'pointfusion' (PointFusion) or 'aggregate' (ICPSLAM).  Returns (SurfelMap, poses (B,L,4,4))."""
    Bn, L, H, W, _ = depth.shape
    dot_th = math.cos(angle_th * math.pi / 180)
    kw = dict(numiters=numiters, damp=damp, dist_thresh=dist_thresh)
    if odom == "gradicp":
        kw.update(lambda_max=lambda_max, B=B, B2=B2, nu=nu)
    smap = SurfelMap()
    out_poses = torch.empty(Bn, L, 4, 4)
    K4 = K[:, 0]
    prev_pose = None
    for s in range(L):
        d = depth[:, s:s + 1]
        c = rgb[:, s:s + 1]
        if s == 0 or odom == "gt":
            pose = torch.eye(4).repeat(Bn, 1, 1) if (poses is None and s == 0) else poses[:, s]
        else:
            at_prev = frame_maps(d, K, prev_pose.unsqueeze(1))
            pose = odometry(smap, at_prev, prev_pose, K4, H, W, odom, dsratio, kw)
        maps = frame_maps(d, K, pose.unsqueeze(1))
        if mode == "pointfusion":
            smap = update_map_fusion(smap, maps, c, pose, K4, dist_th, dot_th, sigma)
        else:
            smap = update_map_aggregate(smap, maps, c)
        prev_pose = pose
        out_poses[:, s] = pose
    return SlamResult(smap, out_poses)
