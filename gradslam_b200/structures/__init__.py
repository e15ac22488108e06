"""Containers of the hot path: batched RGB-D frames and capacity-backed surfel maps."""
from .pointclouds import Pointclouds
from .rgbdimages import RGBDImages
from .utils import pointclouds_from_rgbdimages

__all__ = ["Pointclouds", "RGBDImages", "pointclouds_from_rgbdimages"]
