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
    coordinates normalised by (W,H), weights 1.0 / 0.99).  Default (None): rputil.getKeypoint / getKeypoint_kinect as in
    the reference -- OpenCV SIFT on the CPU, the descriptor-distance augmentation on the GPU (csrc/rp_keypoint.cu), random
    draws from the global numpy state."""
    if keypoint_fn is None:                                                     # rpmodule.py:517-520
        if 'suncg' in dataset or 'matterport' in dataset:
            pts, ptsNorm, ptsW, ptt, pttNorm, pttW = getKeypoint(dataS['rgb'], dataT['rgb'], dataS['feat'], dataT['feat'])
        elif 'scannet' in dataset:
            pts, ptsNorm, ptsW, ptt, pttNorm, pttW = getKeypoint_kinect(dataS['rgb'], dataT['rgb'], dataS['feat'], dataT['feat'],
                                                                        dataS['rgb_full'], dataT['rgb_full'])
        else:
            raise ValueError("unknown dataset %r" % (dataset,))
    else:
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


# ---------------------------------------------------------------------------------------------------------------------
# Module-level fitters with the reference's positional signatures (rpmodule.py:17,60,86,169,212), backed by the
# rp_spectral_irls_solve stage entry.
def horn87_np(src, tgt, weight=None):
    """rpmodule.py:17-58.  src, tgt: [(k),3,n]; weight: [(k),n] -> R [(k),3,3]."""
    src, tgt = np.asarray(src, dtype=np.float64), np.asarray(tgt, dtype=np.float64)
    if src.ndim == 2 and tgt.ndim == 2:
        src, tgt = src[np.newaxis], tgt[np.newaxis]
    assert src.shape[2] == tgt.shape[2]
    k, n = src.shape[0], src.shape[2]
    w = np.ones([k, n]) if weight is None else np.asarray(weight, dtype=np.float64).reshape(k, n)
    sn = src.transpose(0, 2, 1).reshape(-1, 3)
    tn = tgt.transpose(0, 2, 1).reshape(-1, 3)
    zeros = np.zeros_like(sn)
    off = np.arange(k + 1) * n
    T = _solver.default_solver().fit_nodes(zeros, sn, zeros, tn, off, 'horn87', 1.0,
                                           node_w=(np.zeros(k * n), w.reshape(-1)))
    return T[:, :3, :3].copy()


def _fit_rows(allSP, allTP, allSN, allTN, allWP, allWN, mu, method):
    # Stacked rows with identical geometry have identical residuals, hence identical IRLS factors: they can be merged
    # into one node whose base weight is the sum of theirs (the helper stacks every correspondence once per pair).
    geo = np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1, 3) for a in (allSP, allTP, allSN, allTN)], 1)
    uniq, inv = np.unique(geo, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    wp = np.bincount(inv, weights=np.asarray(allWP, dtype=np.float64), minlength=len(uniq))
    wn = np.bincount(inv, weights=np.asarray(allWN, dtype=np.float64), minlength=len(uniq))
    return _solver.default_solver().fit_nodes(uniq[:, 0:3], uniq[:, 6:9], uniq[:, 3:6], uniq[:, 9:12], np.array([0, len(uniq)]),
                                              method, mu, node_w=(wp, wn))[0]


def fit_horn87(allSP, allTP, allSN, allTN, allWP, allWN, mu):
    """rpmodule.py:60-84."""
    return _fit_rows(allSP, allTP, allSN, allTN, allWP, allWN, mu, 'horn87')


def fit_irls(allSP, allTP, allSN, allTN, allWP, allWN, mu):
    """rpmodule.py:169-210."""
    return _fit_rows(allSP, allTP, allSN, allTN, allWP, allWN, mu, 'irls')


def _fit_graph(allSP, allTP, allSN, allTN, allWP, allWN, w, mu, row, col, method):
    """The spectral fitters work on the compact affinity: stacked rows 0..M-1 / M..2M-1 are the first / second
    correspondence of pair i (rpmodule.py:484-489), identified by their flat ids row[i] / col[i] (:495-496)."""
    w = np.asarray(w, dtype=np.float64)
    M = w.shape[0]
    allWP, allWN = np.asarray(allWP, dtype=np.float64), np.asarray(allWN, dtype=np.float64)
    if not (np.array_equal(allWP, np.tile(w, 2)) and np.array_equal(allWN, np.tile(w, 2))):
        raise NotImplementedError("fit_spectral/fit_irls_sm expect allWP == allWN == [w, w], as RelativePoseEstimation_helper "
                                  "builds them (rpmodule.py:488-489)")
    ids = np.concatenate((np.asarray(row), np.asarray(col)))
    uniq, first, inv = np.unique(ids, return_index=True, return_inverse=True)
    for arr in (allSP, allTP, allSN, allTN):
        if not np.array_equal(np.asarray(arr)[first][inv], np.asarray(arr)):
            raise NotImplementedError("stacked rows with the same correspondence id must carry the same geometry")
    rc = np.stack((inv[:M], inv[M:]), 1)
    solver = _solver.default_solver()
    return solver.fit_nodes(np.asarray(allSP)[first], np.asarray(allSN)[first], np.asarray(allTP)[first], np.asarray(allTN)[first],
                            np.array([0, len(uniq)]), method, mu, edges=(np.array([0, M]), rc, w))[0]


def fit_spectral(allSP, allTP, allSN, allTN, allWP, allWN, w_i1i2j1j2, mu, row, col, numFea_s, numFea_t):
    """rpmodule.py:86-167."""
    return _fit_graph(allSP, allTP, allSN, allTN, allWP, allWN, w_i1i2j1j2, mu, row, col, 'spectral')


def fit_irls_sm(allSP, allTP, allSN, allTN, allWP, allWN, w_i1i2j1j2, mu, row, col, numFea_s, numFea_t):
    """rpmodule.py:212-315 (the default method)."""
    return _fit_graph(allSP, allTP, allSN, allTN, allWP, allWN, w_i1i2j1j2, mu, row, col, 'irls+sm')


from ..util import apply_mask          # noqa: E402,F401  (util.apply_mask, util.py:209-232; kept importable from here)


def RelativePoseEstimationViaCompletion(net, data_s, data_t, args, warping_fn=None, keypoint_fn=None):
    """The main algorithm (rpmodule.py:569-662): alternate scan completion (``net`` = SCNet) and pairwise matching.

    args: snumclass, featureDim, outputType, maskMethod, alterStep, dataset, para (per-step sigma arrays),
    representation, completion.  Warping (util.warping, util.py:94-172) and the blend of the completed scans with the
    observed region (:628-634) run on the GPU (relativepose_b200/util.py -> csrc/rp_warp.cu): both warps of a step are one
    batched call and the views never leave the device.  ``warping_fn(view_np [1,8,h,w], R4x4, dataset) -> np [1,8,h,w]``
    overrides the warp (tests); ``keypoint_fn`` is the keypoint stage (SIFT + augmentation are SURVEY 8f row 2)."""
    import copy
    import torch
    from .. import util as _util
    idx_f = 0
    for key, n in (('rgb', 3), ('n', 3), ('d', 1), ('s', args.snumclass)):
        if key in args.outputType:
            idx_f += n
    idx_f_end = idx_f + args.featureDim
    dev = next(net.parameters()).device

    def v(a):
        return torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)

    with torch.no_grad():
        R_hat = np.eye(4)
        full = [torch.cat((v(d['rgb']), v(d['norm']), v(d['depth']).unsqueeze(2)), 2).permute(2, 0, 1).unsqueeze(0)
                for d in (data_s, data_t)]                           # :599-600
        views, masks = [], []
        for c in full:
            vw, m, _geow = apply_mask(c.clone(), args.maskMethod)           # :603-604
            masks.append(m[0, 0])                                    # [h,w] on the device
            views.append(torch.cat((vw, (vw[:, 6:7] != 0).float()), 1))   # :609-612
        view_s, view_t = views
        norm_gt = torch.stack([torch.as_tensor(np.asarray(d['norm'])) for d in (data_s, data_t)]).to(dev)
        depth_gt = torch.stack([torch.as_tensor(np.asarray(d['depth'])) for d in (data_s, data_t)]).to(dev)
        mask2 = torch.stack(masks)
        for alter_ in range(args.alterStep):
            # warp each scan into the other's frame with the current estimate (:616-617); identity -> zeros (util.py:95-96)
            if warping_fn is None:
                w2 = _util.warping_device(torch.cat((view_t, view_s)), np.stack((np.linalg.inv(R_hat), R_hat)), args.dataset)
                view_t2s, view_s2t = w2[0:1], w2[1:2]
            else:
                view_t2s = v(warping_fn(view_t.cpu().numpy(), np.linalg.inv(R_hat), args.dataset)) if np.linalg.norm(R_hat - np.eye(4)) else torch.zeros_like(view_t)
                view_s2t = v(warping_fn(view_s.cpu().numpy(), R_hat, args.dataset)) if np.linalg.norm(R_hat - np.eye(4)) else torch.zeros_like(view_s)
            f = net(torch.cat((torch.cat((view_s, view_t2s), 1), torch.cat((view_t, view_s2t), 1))))   # :619-623
            nrm2, dep2 = _util.blend_completion_device(f, mask2, norm_gt, depth_gt)                    # :628-634
            comp = []
            for k, d in enumerate((data_s, data_t)):
                m = masks[k].cpu().numpy()[:, :, None]
                c = {'normal': nrm2[k].cpu().numpy(), 'depth': dep2[k].cpu().numpy(), 'obs_mask': m.copy(),
                     'rgb': (m * d['rgb'] * 255).astype('uint8'), 'feat': f[k, idx_f:idx_f_end]}   # :636-652
                if 'scannet' in args.dataset:
                    c['rgb_full'] = (d['rgb_full'] * 255).astype('uint8')
                    c['depth_full'] = d['depth_full']
                comp.append(c)
            para_this = copy.copy(args.para)                         # :654-658
            for name in ('sigmaAngle1', 'sigmaAngle2', 'sigmaDist', 'sigmaFeat'):
                setattr(para_this, name, getattr(args.para, name)[alter_])
            R_hat = RelativePoseEstimation(comp[0], comp[1], para_this, args.dataset, args.representation,
                                           doCompletion=args.completion, maskMethod=args.maskMethod, index=None,
                                           keypoint_fn=keypoint_fn)                       # :660
    return R_hat
