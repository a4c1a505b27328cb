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


def getMatchingPrimitive(dataS, dataT, dataset, representation, doCompletion, keypoint_fn=None):
    """rpmodule.py:511-538: keypoints -> 3-D positions / normals / descriptors / observation weights.

    ``keypoint_fn(dataS, dataT, dataset)`` must return the reference's 6-tuple
    ``(pts, ptsNorm, ptsW, ptt, pttNorm, pttW)`` (rputil.getKeypoint / getKeypoint_kinect: pixel coordinates [n,2],
    coordinates normalised by (W,H), weights 1.0 / 0.99).  The reference's own detector is OpenCV-contrib SIFT plus
    an unseeded random augmentation (rputil.py:141-353) and sits outside this repo's parity perimeter
    (SURVEY.md section 8f row 2), so it has to be supplied."""
    if keypoint_fn is None:
        raise NotImplementedError("getMatchingPrimitive needs keypoint_fn (the SIFT keypoint stage of rputil.getKeypoint is "
                                  "outside the B200 hot path; see DESIGN.md section 8)")
    pts, ptsNorm, ptsW, ptt, pttNorm, pttW = keypoint_fn(dataS, dataT, dataset)
    if pts is None or ptt is None or pts.shape[1] < 2 or ptt.shape[1] < 2:
        return None, None, None, None, None, None, None, None
    pts3d, ptsns = getPixel(dataS['depth'], dataS['normal'], pts, dataset=dataset, representation=representation)
    ptt3d, ptsnt = getPixel(dataT['depth'], dataT['normal'], ptt, dataset=dataset, representation=representation)
    dess = interpolate(dataS['feat'], ptsNorm).cpu().numpy().T          # [n,32] float32, as torch_op.npy(...).T
    dest = interpolate(dataT['feat'], pttNorm).cpu().numpy().T
    if not doCompletion:            # keep only keypoints from the observed region (rpmodule.py:534-537)
        ks, kt = ptsW == 1, pttW == 1
        pts3d, ptsns, dess, ptsW = pts3d[:, ks], ptsns[ks], dess[ks], ptsW[ks]
        ptt3d, ptsnt, dest, pttW = ptt3d[:, kt], ptsnt[kt], dest[kt], pttW[kt]
    return pts3d, ptt3d, ptsns, ptsnt, dess, dest, ptsW, pttW


def RelativePoseEstimation(dataS, dataT, para, dataset, representation, maskMethod, doCompletion=True, index=None,
                           keypoint_fn=None):
    """rpmodule.py:540-566: keypoints -> matching primitives -> RelativePoseEstimation_helper."""
    R_hat = np.eye(4)
    prim = getMatchingPrimitive(dataS, dataT, dataset, representation, doCompletion, keypoint_fn)
    pts3d, ptt3d, ptsns, ptsnt, dess, dest, ptsW, pttW = prim
    if pts3d is None or ptt3d is None or pts3d.shape[0] < 2:
        return R_hat
    return RelativePoseEstimation_helper({'pc': pts3d.T, 'normal': ptsns, 'feat': dess, 'weight': ptsW},
                                         {'pc': ptt3d.T, 'normal': ptsnt, 'feat': dest, 'weight': pttW}, para)
