from relativepose_b200.RPModule.rputil import *  # noqa: F401,F403
from relativepose_b200.RPModule.rputil import opts, angular_distance_np, interpolate, getPixel  # noqa: F401
