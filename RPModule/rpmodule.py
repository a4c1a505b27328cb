from relativepose_b200.RPModule.rpmodule import *  # noqa: F401,F403
from relativepose_b200.RPModule.rpmodule import (RelativePoseEstimation_helper, RelativePoseEstimation_batch,  # noqa: F401
                                                  RelativePoseEstimation, getMatchingPrimitive)
