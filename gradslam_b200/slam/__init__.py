from . import fusionutils
from .fusionutils import update_map_aggregate, update_map_fusion
from .icpslam import ICPSLAM
from .pointfusion import PointFusion
