from .cd.chamfer_distance import *  # noqa: F401,F403
from .cd.chamfer_distance import ChamferDistance, ChamferDistanceFunction, chamfer_distance  # noqa: F401
