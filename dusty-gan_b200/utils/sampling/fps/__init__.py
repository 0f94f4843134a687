from .furthest_point_sampling import *  # noqa: F401,F403
from .furthest_point_sampling import downsample_point_clouds, furthest_point_sampling, gather_operation  # noqa: F401
