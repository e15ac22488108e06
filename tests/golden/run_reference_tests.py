"""Runs the REFERENCE's own test files against gradslam_b200 (build container only; /root/reference is read-only and
absent on the GPU box).  `gradslam` and its sub-modules are aliased to this package, the working directory is the
reference root (its tests load tests/data/msrd_b2s3 relative to it) and nothing is written there.

    python tests/golden/run_reference_tests.py [pytest args / test files relative to /root/reference]

Without a GPU only the tests that do not reach a compute kernel can pass (structures, argument checking, views); with a
GPU (and a copy of the reference tests) the same command exercises the kernels through the reference's assertions.
Results of the last run in the build container are recorded in DESIGN.md section 5.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE_ROOT = "/root/reference"
sys.path.insert(0, ROOT)

import gradslam_b200  # noqa: E402

ALIASES = ["", ".structures", ".structures.pointclouds", ".structures.rgbdimages", ".structures.utils", ".structures.structutils", ".geometry",
           ".geometry.projutils", ".geometry.se3utils", ".geometry.geometryutils", ".slam", ".slam.fusionutils",
           ".slam.icpslam", ".slam.pointfusion", ".odometry", ".odometry.icputils", ".odometry.icp", ".odometry.gradicp",
           ".odometry.groundtruth", ".odometry.base"]


def install_aliases():
    for suffix in ALIASES:
        try:
            mod = importlib.import_module("gradslam_b200" + suffix)
        except ImportError as e:  # a sub-module this package does not have
            print("no alias for gradslam%s: %s" % (suffix, e))
            continue
        sys.modules["gradslam" + suffix] = mod
    # visualisation-only third-party imports of the reference's test helpers
    for name in ("open3d", "plotly", "plotly.graph_objects", "plotly.subplots"):
        sys.modules.setdefault(name, types.ModuleType(name))


if __name__ == "__main__":
    import pytest

    install_aliases()
    os.chdir(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    args = sys.argv[1:] or ["tests/structures/test_pointclouds.py", "tests/structures/test_rgbdimages.py",
                            "tests/structures/test_utils.py", "tests/geometry/test_projutils.py"]
    sys.exit(pytest.main(["-p", "no:cacheprovider", "-q", "-rN"] + args))
