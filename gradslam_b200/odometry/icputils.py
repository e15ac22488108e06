"""Point-to-plane ICP building blocks (mirror of gradslam/odometry/icputils.py) — filled in with csrc/gsx_icp.cu."""
__all__ = []


def localize_against_map(pointclouds, live_frame, prev_frame, dsratio, odomprov):
    raise NotImplementedError("ICP odometry kernels are not built yet")
