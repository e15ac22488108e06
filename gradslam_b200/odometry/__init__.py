from .base import OdometryProvider
from .groundtruth import GroundTruthOdometryProvider
from .icp import ICPOdometryProvider
from .gradicp import GradICPOdometryProvider
from . import icputils
