"""CPU emulation of the decision-exact shortcuts of the K2 variant -DGSX_K2_FASTTEST=1 (csrc/gsx_fusion.cu) on the
oracle's data: for every map point inside the frustum of every frame it evaluates, in float32 with the kernel's operation
order, (a) the canonical tests of find_similar_map_points (sqrt(d2) < dist_th, n_frame . n_map > dot_th) and (b) the
shortcut tests (d2 <= d2_max; un-normalised normal test with its guard band), and reports how often the shortcut decides
on its own, how often it falls back to the canonical chain, and that no decision differs.   No GPU needed.

    python scripts/fasttest_guard_check.py [--L 6] [--H 120] [--W 160] [--B 2]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import gsx_oracle as oracle
from gradslam_b200.synthetic import make_sequence

F32 = torch.float32
args = sys.argv[1:]
opts = {"--L": 6, "--B": 2, "--H": 120, "--W": 160}
for k in list(opts):
    if k in args:
        opts[k] = int(args[args.index(k) + 1])
L, B, H, W = opts["--L"], opts["--B"], opts["--H"], opts["--W"]
dist_th, angle_th, sigma = 0.05, 20.0, 0.6
dot_th = math.cos(angle_th * math.pi / 180)
dot_th32 = torch.tensor(dot_th, dtype=F32)


def sqrt_lt_threshold(t):  # as csrc/gsx_thresholds.h
    t = np.float32(t)
    x = np.float32(np.float64(t) * np.float64(t))
    while x > 0 and not (np.sqrt(x, dtype=np.float32) < t):
        x = np.nextafter(x, np.float32(-np.inf), dtype=np.float32)
    while np.sqrt(np.nextafter(x, np.float32(np.inf), dtype=np.float32), dtype=np.float32) < t:
        x = np.nextafter(x, np.float32(np.inf), dtype=np.float32)
    return x


d2_max = torch.tensor(float(sqrt_lt_threshold(dist_th)), dtype=F32)
rgb, depth, K, poses = make_sequence(B, L, H, W, seed=0)
smap = oracle.SurfelMap()
tot = dict(active=0, close=0, fast=0, fallback=0, mismatch_dist=0, mismatch_normal=0, max_err=0.0)
for s in range(L):
    maps = oracle.frame_maps(depth[:, s:s + 1], K, poses[:, s:s + 1])
    if smap.has_points:
        table = oracle.find_active_map_points(smap, poses[:, s], K[:, 0], H, W)
        b, n, h, w = table.unbind(1)
        pts, nrm, _, _ = smap.padded()
        mp, mn = pts[b, n], nrm[b, n]
        gv, gn = maps["gvertex"][:, 0][b, h, w], maps["gnormal"][:, 0][b, h, w]
        d = gv - mp
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        close_exact = oracle._sqrt32(d2) < torch.tensor(dist_th, dtype=F32)
        close_fast = d2 <= d2_max
        dot_exact = (gn[:, 0] * mn[:, 0] + gn[:, 1] * mn[:, 1]) + gn[:, 2] * mn[:, 2]
        sim_exact = dot_exact > dot_th32
        # un-normalised cross product with the kernel's neighbour rule (last column / row re-use their neighbour's)
        vert = maps["vertex"][:, 0]
        wa, ha = w.clamp(max=W - 2), h.clamp(max=H - 2)
        dh = vert[b, h, wa + 1] - vert[b, h, wa]
        dv = vert[b, ha + 1, w] - vert[b, ha, w]
        cx = dh[:, 1] * dv[:, 2] - dh[:, 2] * dv[:, 1]
        cy = dh[:, 2] * dv[:, 0] - dh[:, 0] * dv[:, 2]
        cz = dh[:, 0] * dv[:, 1] - dh[:, 1] * dv[:, 0]
        c2 = (cx * cx + cy * cy) + cz * cz
        R = poses[:, s][b][:, :3, :3]
        rc = [oracle._dot3(R[:, i, 0], R[:, i, 1], R[:, i, 2], cx, cy, cz) for i in range(3)]
        approx = ((rc[0] * mn[:, 0] + rc[1] * mn[:, 1]) + rc[2] * mn[:, 2]) * (1.0 / torch.sqrt(c2)).to(F32)
        guard = 1e-4 * (1.0 + (mn[:, 0].abs() + mn[:, 1].abs()) + mn[:, 2].abs())
        valid = maps["valid"][:, 0][b, h, w]
        usable = valid & (c2 > 1e-30) & (c2 < 1e30)
        decided = usable & ((approx - dot_th32).abs() > guard)
        sim_fast = torch.where(decided, approx > dot_th32, sim_exact)  # inside the band the kernel runs the canonical chain
        cl = close_exact
        tot["active"] += table.shape[0]
        tot["close"] += int(cl.sum())
        tot["fast"] += int((cl & decided).sum())
        tot["fallback"] += int((cl & ~decided).sum())
        if bool(usable.any()):
            tot["max_err"] = max(tot["max_err"], float((approx - dot_exact)[usable].abs().max()))
        tot["mismatch_dist"] += int((close_exact != close_fast).sum())
        tot["mismatch_normal"] += int((cl & (sim_fast != sim_exact)).sum())
    smap = oracle.update_map_fusion(smap, maps, rgb[:, s:s + 1], poses[:, s], K[:, 0], dist_th, dot_th, sigma)
    print("frame %d: map %s" % (s, smap.counts()), flush=True)
print(tot)
print("fallback rate among close points: %.4f %%; distance-test mismatches %d; normal-test mismatches %d; "
      "max |shortcut - canonical| of the normal product %.2e (guard band >= 1e-4)" % (
          100.0 * tot["fallback"] / max(1, tot["close"]), tot["mismatch_dist"], tot["mismatch_normal"], tot["max_err"]))
assert tot["mismatch_dist"] == 0 and tot["mismatch_normal"] == 0
