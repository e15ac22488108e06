"""Ground-truth odometry (mirror of gradslam/odometry/groundtruth.py:9-75): T = T1^{-1} T2."""
from ..geometry.geometryutils import relative_transformation
from ..structures.rgbdimages import RGBDImages
from .base import OdometryProvider

__all__ = ["GroundTruthOdometryProvider"]


class GroundTruthOdometryProvider(OdometryProvider):
    def __init__(self):
        pass

    def provide(self, rgbdimages1: RGBDImages, rgbdimages2: RGBDImages):
        """Relative pose between two sequence-length-1 batches that carry poses.  Returns (B, 1, 4, 4)."""
        for i, r in ((1, rgbdimages1), (2, rgbdimages2)):
            if not isinstance(r, RGBDImages):
                raise TypeError("Expected rgbdimages{0} to be of type gradslam.RGBDImages. Got {1}.".format(i, type(r)))
        for i, r in ((1, rgbdimages1), (2, rgbdimages2)):
            if r.poses is None:
                raise ValueError("Input {0} (rgbdimages{0}) missing poses. Poses must be provided if using "
                                 "GroundTruthOdometryProvider".format(i))
        for i, r in ((1, rgbdimages1), (2, rgbdimages2)):
            if r.shape[1] != 1:
                raise ValueError("Sequence length of rgbdimages{0} must be 1, but was {1}.".format(i, r.shape[1]))
        if rgbdimages1.shape[0] != rgbdimages2.shape[0]:
            raise ValueError("Batch size of rgbdimages1 and rgbdimages2 should be equal ({0} != {1})".format(
                rgbdimages1.shape[0], rgbdimages2.shape[0]))
        B, L = rgbdimages1.shape[:2]
        return relative_transformation(rgbdimages1.poses.view(-1, 4, 4), rgbdimages2.poses.view(-1, 4, 4),
                                       orthogonal_rotations=False).view(B, L, 4, 4)
