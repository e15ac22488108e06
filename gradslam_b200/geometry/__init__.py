from .projutils import *
from .se3utils import *
from .geometryutils import *
