from .pointclouds import Pointclouds
from .rgbdimages import RGBDImages
from .utils import pointclouds_from_rgbdimages
