"""Frame -> cloud conversion (mirror of gradslam/structures/utils.py:7-57)."""
import torch

from .pointclouds import Pointclouds
from .rgbdimages import RGBDImages

__all__ = ["pointclouds_from_rgbdimages"]


def pointclouds_from_rgbdimages(rgbdimages: RGBDImages, *, global_coordinates: bool = True,
                                filter_missing_depths: bool = True) -> Pointclouds:
    """Converts a sequence-length-1 RGBDImages batch into Pointclouds (points, normals, colors).

    With `filter_missing_depths` the valid pixels of every element are compacted in row-major order by the
    stable-append kernel (the same K4 kernel PointFusion uses, run with no matches)."""
    if not isinstance(rgbdimages, RGBDImages):
        raise TypeError("Expected rgbdimages to be of type gradslam.RGBDImages. Got {0}.".format(type(rgbdimages)))
    if not rgbdimages.shape[1] == 1:
        raise ValueError("Expected rgbdimages to have sequence length of 1. Got {0}.".format(rgbdimages.shape[1]))
    B = rgbdimages.shape[0]
    rgbdimages = rgbdimages.to_channels_last()
    if filter_missing_depths:
        from ..slam.fusionutils import _append_valid_pixels

        return _append_valid_pixels(Pointclouds(device=rgbdimages.device), rgbdimages, global_coordinates)
    vmap = rgbdimages.global_vertex_map if global_coordinates else rgbdimages.vertex_map
    nmap = rgbdimages.global_normal_map if global_coordinates else rgbdimages.normal_map
    return Pointclouds(points=vmap.reshape(B, -1, 3).contiguous(), normals=nmap.reshape(B, -1, 3).contiguous(),
                       colors=rgbdimages.rgb_image.reshape(B, -1, 3).contiguous())
