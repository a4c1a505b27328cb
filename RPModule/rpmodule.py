from relativepose_b200.RPModule.rpmodule import *  # noqa: F401,F403
from relativepose_b200.RPModule.rpmodule import (RelativePoseEstimation_helper, RelativePoseEstimation_batch,  # noqa: F401
                                                  RelativePoseEstimation, RelativePoseEstimationViaCompletion,
                                                  getMatchingPrimitive, apply_mask,
                                                  horn87_np, fit_horn87, fit_irls, fit_spectral, fit_irls_sm)
