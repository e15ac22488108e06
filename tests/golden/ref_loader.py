"""Import the UNMODIFIED gradslam reference from /root/reference with in-memory shims.

TEST INFRASTRUCTURE ONLY.  This module exists so that `make_golden.py` (run in
the build container, where /root/reference is mounted) can execute the real
reference and freeze its outputs as fixtures under tests/golden/.  Nothing in
the product package, in `-m gpu` tests, in smoke() or in bench.py imports it:
/root/reference does not exist on the GPU box.

Four third-party modules the reference imports are not installable offline
(SURVEY.md §8c):

  * open3d, plotly          -- visualisation only; empty stubs.
  * kornia.geometry.linalg  -- two 4x4 helpers (compose / inverse rigid
                               transform); restated below.
  * chamferdist.chamfer     -- knn_points (pinned chamferdist==1.0.0, call site
                               gradslam/odometry/icputils.py:200); restated as
                               an exact brute-force squared-L2 1-NN with
                               lowest-index tie-break.
"""
import importlib
import sys
import types
from collections import namedtuple

import torch

REFERENCE_ROOT = "/root/reference"


def _stub(name):
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


def _install_shims():
    if "chamferdist" in sys.modules and hasattr(sys.modules["chamferdist"], "_gsx_shim"):
        return
    # --- open3d / plotly: never called on the hot path --------------------------------
    o3d = _stub("open3d")
    o3d.geometry = types.SimpleNamespace()
    o3d.utility = types.SimpleNamespace()
    plotly = _stub("plotly")
    go = _stub("plotly.graph_objects")
    sub = _stub("plotly.subplots")
    sub.make_subplots = lambda *a, **k: None
    plotly.graph_objects = go
    plotly.subplots = sub

    # --- kornia.geometry.linalg ---------------------------------------------------------
    kornia = _stub("kornia")
    kgeo = _stub("kornia.geometry")
    klin = _stub("kornia.geometry.linalg")

    def compose_transformations(trans_01, trans_12):
        r01, t01 = trans_01[..., :3, :3], trans_01[..., :3, 3:]
        r12, t12 = trans_12[..., :3, :3], trans_12[..., :3, 3:]
        r02 = torch.matmul(r01, r12)
        t02 = torch.matmul(r01, t12) + t01
        out = torch.zeros_like(trans_01)
        out[..., :3, :3] = r02
        out[..., :3, 3:] = t02
        out[..., 3, 3] = 1.0
        return out

    def inverse_transformation(trans_12):
        r12, t12 = trans_12[..., :3, :3], trans_12[..., :3, 3:]
        r21 = r12.transpose(-1, -2)
        t21 = torch.matmul(-r21, t12)
        out = torch.zeros_like(trans_12)
        out[..., :3, :3] = r21
        out[..., :3, 3:] = t21
        out[..., 3, 3] = 1.0
        return out

    klin.compose_transformations = compose_transformations
    klin.inverse_transformation = inverse_transformation
    kgeo.linalg = klin
    kornia.geometry = kgeo

    # --- chamferdist.chamfer.knn_points --------------------------------------------------
    chamferdist = _stub("chamferdist")
    chamfer = _stub("chamferdist.chamfer")
    _KNN = namedtuple("KNN", "dists idx knn")

    def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1,
                   return_nn=False, return_sorted=True):
        assert K == 1 and p1.shape[0] == 1 and p2.shape[0] == 1
        a, b = p1[0], p2[0]
        dists = torch.empty(a.shape[0], dtype=a.dtype, device=a.device)
        idx = torch.empty(a.shape[0], dtype=torch.int64, device=a.device)
        chunk = 2048
        for s in range(0, a.shape[0], chunk):
            diff = a[s:s + chunk, None, :] - b[None, :, :]
            d = (diff * diff).sum(-1)
            m, i = d.min(dim=1)  # first minimum == lowest index on CPU
            dists[s:s + chunk] = m
            idx[s:s + chunk] = i
        return _KNN(dists.view(1, -1, 1), idx.view(1, -1, 1), None)

    chamfer.knn_points = knn_points
    chamferdist.chamfer = chamfer
    chamferdist._gsx_shim = True


def load_reference():
    """Returns the reference `gradslam` module (imported from /root/reference)."""
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("gradslam")
