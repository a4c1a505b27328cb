"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference RPModule solver.

This is the *oracle*: a numpy/scipy restatement of
``RPModule/rpmodule.py:17-508`` (reference @ 2e9fdf5) used as the checker in
``tests/``, in ``__graft_entry__.smoke()`` and as ``bench.py``'s CPU baseline
(``cpu_baseline.kind == "port"``).  Nothing in the product path
(``relativepose_b200/``) may import it; the product fails loudly when its CUDA
library is missing.

Pinning: the reference has no tests and no golden vectors,
so the oracle is pinned against *outputs of the reference itself*, executed in
the build container through ``oracle/ref_loader.py`` and frozen as
``tests/golden/rp_golden_*.npz`` by ``tests/golden/make_golden.py``.
``tests/test_oracle_golden.py`` checks this file against those vectors (top-k
sets and surviving-pair masks exactly, 4x4 poses to 1e-12) and, when
/root/reference is present, against the live reference.

Each function cites the reference lines it restates.  The arithmetic keeps the
reference's data types and library calls (float32 descriptor distances summed
by numpy, float64 afterwards, ``scipy.sparse.linalg.eigs`` for the leading
eigenvector, ``numpy.linalg.eig`` inside Horn's method) so that its timing is a
fair stand-in for the reference's CPU path.
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import eigs

FEAT_SCALING = 100      # rpmodule.py:327
OBS_W = 1.2             # rpmodule.py:328
UNOBSERVED_DAMP = 0.6   # rpmodule.py:467
OFFSET = 50             # rpmodule.py:103,231
NUM_ALTER = 5           # rpmodule.py:102,229
NUM_REWEIGHT = 5        # rpmodule.py:181,228
EPS = 1e-12             # rpmodule.py:71,104,183,232

STATUS_OK = 0
STATUS_FEW_KEYPOINTS = 1      # rpmodule.py:346-348
STATUS_FEW_CORRES = 2         # rpmodule.py:377-379
STATUS_FEW_DIST = 3           # rpmodule.py:406-408
STATUS_FEW_ANGLE = 4          # rpmodule.py:440-443
STATUS_ZERO_WEIGHT = 5        # rpmodule.py:469-472


class Params(object):
    """Same fields and defaults as ``rputil.opts`` (RPModule/rputil.py:11-22)."""

    def __init__(self, sigmaAngle1=0.523 / 2, sigmaAngle2=0.523 / 2, sigmaDist=0.08 / 2, sigmaFeat=0.01):
        self.distThre = 0.08
        self.distSepThre = 1.5 * 0.08
        self.angleThre = 45 / 180. * np.pi
        self.sigmaAngle1 = sigmaAngle1
        self.sigmaAngle2 = sigmaAngle2
        self.sigmaDist = sigmaDist
        self.sigmaFeat = sigmaFeat
        self.mu = 0.3
        self.topK = 5
        self.method = 'irls+sm'


# --------------------------------------------------------------------------- Horn
def horn_rotation(src, tgt, weight):
    """Weighted Horn-87 rotation, one problem.  rpmodule.py:17-58.

    src, tgt: [3, n]; weight: [n].  Returns R [3,3]."""
    M = np.matmul(src[None], (tgt * weight[None, :])[None].transpose(0, 2, 1))[0]   # :39-43
    N = np.array([                                                                    # :46-49
        [M[0, 0] + M[1, 1] + M[2, 2], M[1, 2] - M[2, 1], M[2, 0] - M[0, 2], M[0, 1] - M[1, 0]],
        [M[1, 2] - M[2, 1], M[0, 0] - M[1, 1] - M[2, 2], M[0, 1] + M[1, 0], M[0, 2] + M[2, 0]],
        [M[2, 0] - M[0, 2], M[0, 1] + M[1, 0], M[1, 1] - M[0, 0] - M[2, 2], M[1, 2] + M[2, 1]],
        [M[0, 1] - M[1, 0], M[2, 0] + M[0, 2], M[1, 2] + M[2, 1], M[2, 2] - M[0, 0] - M[1, 1]]])
    vals, vecs = np.linalg.eig(N)                                                     # :50
    q = vecs[:, vals.argmax()]                                                        # :51-53
    a, b, c, d = q
    return np.array([                                                                 # :54-56
        [a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
        [2 * (c * b + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
        [2 * (d * b - a * c), 2 * (d * c + a * b), a * a - b * b - c * c + d * d]])


class _Rows(object):
    """The stacked per-row arrays the fitters work on (rpmodule.py:484-489)."""

    def __init__(self, SP, TP, SN, TN, w_pair):
        self.SP, self.TP, self.SN, self.TN = SP, TP, SN, TN
        self.w_pair = w_pair

    def centre(self, wp):
        """Weighted centroids and centred positions.  rpmodule.py:72-75 et al."""
        sm = (self.SP * wp[:, None]).sum(0) / (wp.sum() + EPS)
        tm = (self.TP * wp[:, None]).sum(0) / (wp.sum() + EPS)
        return sm, tm, self.SP - sm, self.TP - tm

    def horn(self, SPc, TPc, allW, sm, tm):
        """One weighted Horn fit + translation.  rpmodule.py:76-80 et al."""
        S = np.concatenate((SPc, self.SN)).T
        T = np.concatenate((TPc, self.TN)).T
        R = horn_rotation(S, T, allW)
        t = -np.matmul(R, sm) + tm
        return R, t

    def residuals(self, R, SPc, TPc, mu):
        """Per-row residuals.  rpmodule.py:202-203,252-253,304-305."""
        rp = mu * np.power(np.matmul(R, SPc.T) - TPc.T, 2).sum(0)
        rn = np.power(np.matmul(R, self.SN.T) - self.TN.T, 2).sum(0)
        return rp, rn


def _pose(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def _irls_rounds(rows, allW, mu, trace=None):
    """NUM_REWEIGHT rounds of {centre, Horn, residual, reweight}.  rpmodule.py:185-205,236-255,287-307."""
    half = len(allW) // 2
    for _ in range(NUM_REWEIGHT):
        wp = allW[:half]
        sm, tm, SPc, TPc = rows.centre(wp)
        R, t = rows.horn(SPc, TPc, allW, sm, tm)
        rp, rn = rows.residuals(R, SPc, TPc, mu)
        allW = allW * 1.0 / (1.0 + np.concatenate((rp, rn)))      # resSigma = 1
    return R, t, SPc, TPc, allW


# 'LM' (largest magnitude) is what the reference's bare ``eigs(A, k=1)`` asks ARPACK for.  On a bipartite affinity
# (lambda_max = |lambda_min| exactly) ARPACK may return either end of the spectrum, platform dependent; tests of that
# ill-posed situation switch to 'LR' (largest real part = the Perron vector, which the CUDA path always returns).
EIG_WHICH = 'LM'


def _leading_eigvec(a_pair, row, col, dim):
    """Sparse symmetric affinity, leading eigenvector, unit norm.  rpmodule.py:131-136,270-276."""
    A = sp.csc_matrix((a_pair, (row, col)), shape=(dim, dim))
    A = A + A.T
    _, u = eigs(A, k=1, which=EIG_WHICH)
    u = u.real
    u /= np.linalg.norm(u)
    return u


def fit_horn87(rows, mu):
    """rpmodule.py:60-84."""
    allWP = np.concatenate((rows.w_pair, rows.w_pair))
    sm, tm, SPc, TPc = rows.centre(allWP)
    allW = np.concatenate((allWP * mu, allWP))
    R, t = rows.horn(SPc, TPc, allW, sm, tm)
    return _pose(R, t)


def fit_irls(rows, mu):
    """rpmodule.py:169-210."""
    allWP = np.concatenate((rows.w_pair, rows.w_pair))
    allW = np.concatenate((allWP * mu, allWP))
    R, t, _, _, _ = _irls_rounds(rows, allW, mu)
    return _pose(R, t)


def fit_spectral(rows, mu, row, col, dim, trace=None):
    """rpmodule.py:86-167."""
    w = rows.w_pair
    allWP = np.concatenate((w, w))
    sm, tm, SPc, TPc = rows.centre(allWP)
    allW = np.concatenate((allWP * mu, allWP))
    R, t = rows.horn(SPc, TPc, allW, sm, tm)
    for _ in range(NUM_ALTER):
        rp, rn = rows.residuals(R, SPc, TPc, mu)
        a = allWP * (OFFSET - (rp + rn))                          # :126 (current allWP)
        a[a < 0] = 0
        a = a.reshape(2, -1).sum(0)
        u = _leading_eigvec(a, row, col, dim)
        x = (u[row] * u[col]).squeeze()
        x[x < 0] = 0
        x *= w
        if trace is not None:
            trace.setdefault('x', []).append(x.copy())
        allW = np.tile(x, 4)
        allW[:len(allW) // 2] *= mu
        allWP = allW[:len(allW) // 2]
        sm, tm, SPc, TPc = rows.centre(allWP)
        R, t = rows.horn(SPc, TPc, allW, sm, tm)
    return _pose(R, t)


def fit_irls_sm(rows, mu, row, col, dim, trace=None):
    """The default method.  rpmodule.py:212-315."""
    w = rows.w_pair
    allWP = np.concatenate((w, w))
    allW = np.concatenate((allWP * mu, allWP))
    R, t, SPc, TPc, _ = _irls_rounds(rows, allW, mu)
    if trace is not None:
        trace['T_init'] = _pose(R, t)
    for _ in range(NUM_ALTER):
        rp, rn = rows.residuals(R, SPc, TPc, mu)
        a = np.tile(w, 2) * (OFFSET - (rp + rn))                  # :265
        a[a < 0] = 0
        a = a.reshape(2, -1).sum(0)
        u = _leading_eigvec(a, row, col, dim)
        x = (u[row] * u[col]).squeeze()
        x[x < 0] = 0
        x *= w
        allW = np.tile(x, 4)
        allW[:len(allW) // 2] *= mu
        R, t, SPc, TPc, _ = _irls_rounds(rows, allW, mu)
        if trace is not None:
            trace.setdefault('x', []).append(x.copy())
            trace.setdefault('T_alter', []).append(_pose(R, t))
    return _pose(R, t)


# ------------------------------------------------------------------ front end
def descriptor_weights(feat_s, feat_t, w_s, w_t, sigma_feat):
    """Row-normalised soft matches ``wij`` [n_s,n_t] float64.  rpmodule.py:342-363."""
    ds = feat_s / FEAT_SCALING
    dt = feat_t / FEAT_SCALING
    both = w_s[:, None] * w_t[None, :]
    dij = np.power(ds[:, None, :] - dt[None, :, :], 2).sum(2)     # float32 for float32 features
    sig = np.ones(both.shape) * sigma_feat
    sig[both == 1] = sigma_feat / OBS_W
    wij = np.exp(np.divide(-dij, 2 * np.power(sig / 5, 2)))
    nm = np.linalg.norm(wij, axis=1, keepdims=True)
    dead = (nm == 0)
    nm[dead] = 1
    wij /= nm
    wij[dead.squeeze(1), :] = 0
    return wij, dij


def topk_candidates(wij, topk):
    """``corres`` [2, n_s*K] int.  rpmodule.py:368-375."""
    K = min(topk, wij.shape[1] - 1)
    top = np.argpartition(-wij, K, axis=1)[:, :K]
    n_s = wij.shape[0]
    corres = np.zeros([2, n_s * K], dtype=np.int64)
    corres[0] = np.arange(n_s).repeat(K)
    corres[1] = top.flatten()
    return corres, K


def solve_pair(src, tgt, para, trace=None):
    """Restatement of ``RelativePoseEstimation_helper`` (rpmodule.py:317-508).

    src/tgt: dicts 'pc' [k,3], 'normal' [k,3], 'feat' [k,32], 'weight' [k].
    Returns the 4x4 float64 pose; degenerate inputs give identity.  When
    ``trace`` is a dict the stage-boundary quantities are stored in it
    (``status``, ``topk`` [n_s,K] sorted, ``corres``, ``pairs`` [M,2] (first,second
    correspondence index), ``w`` [M], ...)."""
    Ps, Pt = src['pc'], tgt['pc']
    Ns, Nt = src['normal'], tgt['normal']
    ws, wt = src['weight'], tgt['weight']

    def out(status, T=None):
        if trace is not None:
            trace['status'] = status
        return np.eye(4) if T is None else T

    if Ps.shape[0] < 3 or Pt.shape[0] < 3:
        return out(STATUS_FEW_KEYPOINTS)
    n_s, n_t = Ps.shape[0], Pt.shape[0]

    wij, dij = descriptor_weights(src['feat'], tgt['feat'], ws, wt, para.sigmaFeat)
    corres, K = topk_candidates(wij, para.topK)
    ncor = corres.shape[1]
    if trace is not None:
        trace['dij'] = dij
        trace['wij'] = wij
        trace['topk'] = np.sort(corres[1].reshape(n_s, K), axis=1)
        trace['corres'] = corres
    if ncor < 3:
        return out(STATUS_FEW_CORRES)

    # all unordered pairs (first < second), row-major.  rpmodule.py:382-386
    first, second = np.triu_indices(ncor, k=1)
    i1, j1 = corres[0, first], corres[1, first]
    i2, j2 = corres[0, second], corres[1, second]

    # distance consistency.  rpmodule.py:389-404
    dis_s = np.linalg.norm(Ps[i1] - Ps[i2], axis=1)
    dis_t = np.linalg.norm(Pt[j1] - Pt[j2], axis=1)
    dd = np.power(dis_s - dis_t, 2)
    keep = np.logical_and(dd < np.power(para.distThre, 2),
                          np.minimum(dis_s, dis_t) > 1.5 * np.power(para.distSepThre, 2))
    if trace is not None:
        trace['n_dist'] = int(keep.sum())
    if keep.sum() < 3:
        return out(STATUS_FEW_DIST)
    first, second, dd = first[keep], second[keep], dd[keep]
    i1, j1, i2, j2 = i1[keep], j1[keep], i2[keep], j2[keep]

    # angle consistency.  rpmodule.py:424-436
    e1 = Ps[i1] - Ps[i2]
    e2 = Pt[j1] - Pt[j2]
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 /= np.linalg.norm(e2, axis=1, keepdims=True)

    def ang(a, b):
        return np.arccos((a * b).sum(1).clip(-1, 1))

    alpha = np.power(ang(Ns[i1], Ns[i2]) - ang(Nt[j1], Nt[j2]), 2)
    beta = np.power(ang(Ns[i1], e1) - ang(Nt[j1], e2), 2)
    gamma = np.power(ang(Ns[i2], e1) - ang(Nt[j2], e2), 2)
    lim = np.power(para.angleThre, 2)
    keep = np.logical_and.reduce((alpha < lim, beta < lim, gamma < lim))
    if trace is not None:
        trace['n_angle'] = int(keep.sum())
    if keep.sum() < 3:
        return out(STATUS_FEW_ANGLE)
    first, second = first[keep], second[keep]
    i1, j1, i2, j2 = i1[keep], j1[keep], i2[keep], j2[keep]
    dd, alpha, beta, gamma = dd[keep], alpha[keep], beta[keep], gamma[keep]

    # pair weight.  rpmodule.py:453-467
    w = wij[i1, j1] * wij[i2, j2] * np.exp(-dd / (2 * para.sigmaDist ** 2)
                                            - alpha / (2 * para.sigmaAngle1 ** 2)
                                            - beta / (2 * para.sigmaAngle2 ** 2)
                                            - gamma / (2 * para.sigmaAngle2 ** 2))
    seen = ws[i1] * ws[i2] * wt[j1] * wt[j2]
    w[seen != 1] *= UNOBSERVED_DAMP
    if trace is not None:
        trace['pairs'] = np.stack((first, second), 1)
        trace['w'] = w.copy()
    if (w != 0).sum() < 1:
        return out(STATUS_ZERO_WEIGHT)

    rows = _Rows(np.concatenate((Ps[i1], Ps[i2])), np.concatenate((Pt[j1], Pt[j2])),
                 np.concatenate((Ns[i1], Ns[i2])), np.concatenate((Nt[j1], Nt[j2])), w)
    if para.method == 'horn87':
        T = fit_horn87(rows, para.mu)
    elif para.method == 'irls':
        T = fit_irls(rows, para.mu)
    elif para.method in ('spectral', 'irls+sm'):
        row = i1 * n_t + j1
        col = i2 * n_t + j2
        fit = fit_spectral if para.method == 'spectral' else fit_irls_sm
        T = fit(rows, para.mu, row, col, n_s * n_t, trace)
    else:
        raise Exception("unknown method!")                       # rpmodule.py:507-508
    return out(STATUS_OK, T)


def solve_batch(records, para):
    """[B,4,4] poses for a list of primitive-cache records
    (trainRelativePoseModuleRecFD.py:207-208)."""
    out = np.zeros([len(records), 4, 4])
    for b, r in enumerate(records):
        s = {'pc': r['pc_src'], 'normal': r['normal_src'], 'feat': r['feat_src'], 'weight': r['weight_src']}
        t = {'pc': r['pc_tgt'], 'normal': r['normal_tgt'], 'feat': r['feat_tgt'], 'weight': r['weight_tgt']}
        out[b] = solve_pair(s, t, para)
    return out
