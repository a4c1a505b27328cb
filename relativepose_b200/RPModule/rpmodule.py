"""Drop-in for the reference's RPModule/rpmodule.py solver entry points, backed by the CUDA library.

``RelativePoseEstimation_helper(dataS, dataT, para)`` keeps the reference's
signature, input layout and behaviour (RPModule/rpmodule.py:317-508): pure,
synchronous, returns a 4x4 float64 pose, identity on the five degenerate exits,
raises ``Exception("unknown method!")`` for an unknown ``para.method``.
``RelativePoseEstimation_batch`` is the batched form (the single-pair helper is
B=1); its record layout is the reference's primitive cache
(trainRelativePoseModuleRecFD.py:207-208).

There is no CPU fallback: without the CUDA library / a GPU these functions raise.
"""
import logging

import numpy as np

from .rputil import *          # noqa: F401,F403  (the reference does the same: rpmodule.py:9)
from .. import solver as _solver

logger = logging.getLogger(__name__)


def _record(dataS, dataT):
    return {'pc_src': dataS['pc'], 'normal_src': dataS['normal'], 'feat_src': dataS['feat'], 'weight_src': dataS['weight'],
            'pc_tgt': dataT['pc'], 'normal_tgt': dataT['normal'], 'feat_tgt': dataT['feat'], 'weight_tgt': dataT['weight']}


def RelativePoseEstimation_batch(records, para, device=None, return_stats=False):
    """Poses [B,4,4] float64 for a list of primitive-cache records; one fused GPU launch."""
    if para.method not in ('horn87', 'spectral', 'irls', 'irls+sm'):
        raise Exception("unknown method!")                      # rpmodule.py:507-508
    if len(records) == 0:
        return np.zeros([0, 4, 4])
    return _solver.default_solver(device).solve_records(records, para, return_stats=return_stats)


def RelativePoseEstimation_helper(dataS, dataT, para):
    """Given two sets of keypoints ('pc' [k,3], 'normal' [k,3], 'feat' [k,32], 'weight' [k]) estimate the
    relative pose; ``para`` is an ``opts`` (rputil.py).  Reference: rpmodule.py:317-508."""
    if np.asarray(dataS['pc']).shape[0] < 3 or np.asarray(dataT['pc']).shape[0] < 3:
        logger.info("stage-1: not enough!")                    # rpmodule.py:346-348 (before the method dispatch)
        return np.eye(4)
    return RelativePoseEstimation_batch([_record(dataS, dataT)], para)[0]
