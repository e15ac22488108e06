from . import fusionutils
from .icpslam import ICPSLAM
from .pointfusion import PointFusion
