"""The finite-difference parameter trainer's objective as a batched call.

Reference: trainRelativePoseModuleRecFD.py -- ``objective(para)`` (:215-233) solves every cached pair of matching
primitives with RelativePoseEstimation_helper in a Python loop (~20 x |val| solves per outer iteration, :246-291).
Here the primitive cache is uploaded once (``Objective(primitives)``), every evaluation is one fused launch of
rp_solve_batch over all pairs, and only the [B,4,4] poses come back.  The record layout is the reference's cache
(:207-212): a list of dicts {'pc_src','normal_src','feat_src','weight_src','pc_tgt','normal_tgt','feat_tgt','weight_tgt',
'R_gt'} stored with ``np.save`` (pickled object array).

The loss / angular-distance reductions follow the reference's numpy expressions (accumulated pair by pair in the same
order), so a given set of poses yields the same numbers.
"""
import numpy as np

from . import solver as _solver
from .RPModule.rputil import angular_distance_np, opts


def save_primitives(path, primitives):
    """trainRelativePoseModuleRecFD.py:212 (``np.save(primitive_file, primitives)``)."""
    np.save(path, np.array(list(primitives), dtype=object), allow_pickle=True)


def load_primitives(path):
    """trainRelativePoseModuleRecFD.py:113 (``np.load(primitive_file)``; an object array of dicts)."""
    return list(np.load(path, allow_pickle=True))


class Objective(object):
    """``objective(para) -> (loss, ad)`` of trainRelativePoseModuleRecFD.py:215-233 over a fixed primitive cache."""

    def __init__(self, primitives, device=None):
        self.primitives = list(primitives)
        self.solver = _solver.default_solver(device)
        self.packed = _solver.PackedBatch(self.primitives)
        self.dbatch = self.packed.to_device(self.solver.device)
        self.R_gt = np.stack([np.asarray(p['R_gt'], dtype=np.float64) for p in self.primitives]) if self.primitives else np.zeros((0, 4, 4))
        self.evaluations = 0

    def poses(self, para):
        """[B,4,4] float64 poses for this parameter set (one launch over the resident batch)."""
        if para.method not in ('horn87', 'spectral', 'irls', 'irls+sm'):
            raise Exception("unknown method!")
        if not self.primitives:
            return np.zeros((0, 4, 4))
        T, status, _ = self.solver.solve_device_checked(self.dbatch, [_solver.params_from_opts(para)])
        self.evaluations += 1
        return T.cpu().numpy()

    def __call__(self, para):
        T = self.poses(para)
        loss, ad, count = 0, 0, 0
        for i in range(len(self.primitives)):                       # :219-229, same accumulation order
            R_hat, R_gt = T[i], self.R_gt[i]
            loss += np.power(R_hat[:3, :3] - R_gt[:3, :3], 2).sum()
            ad += angular_distance_np(R_hat[:3, :3].reshape(1, 3, 3), R_gt[:3, :3].reshape(1, 3, 3))[0]
            count += 1
        loss /= count
        ad /= count
        return loss, ad


def fd_step(objective, cur, rng, n_probe=10):
    """One outer iteration of the finite-difference descent (trainRelativePoseModuleRecFD.py:246-297).

    cur: [sigmaAngle1, sigmaAngle2, sigmaDist, sigmaFeat]; rng: numpy RandomState (the reference uses the global one,
    ``np.random.uniform(np.zeros([4]))``, :252).  Returns (new_cur, loss_best, ad_best, found_descent)."""
    cur = np.asarray(cur, dtype=np.float64)
    eps = np.zeros([n_probe, 4])
    losses, ads = np.zeros([n_probe]), np.zeros([n_probe])

    def para_of(v):
        p = opts()
        p.sigmaAngle1, p.sigmaAngle2, p.sigmaDist, p.sigmaFeat = [float(x) for x in v]
        return p

    for j in range(n_probe):
        if j >= 1:
            eps[j, :] = (rng.uniform(np.zeros([4])) - 0.5) / 5
        losses[j], ads[j] = objective(para_of(cur * (1 + eps[j])))
    grad = np.linalg.lstsq(eps[1:, :], losses[1:] - losses[0], rcond=-1)[0]
    scale = max(np.abs(grad / cur))
    grad = grad / scale
    alpha, lr = 1, 1
    loss_best, ad_best = losses[0], ads[0]
    found = False
    for j in range(10):
        cand = cur * (1 + -lr * grad * alpha)
        loss_c, ad_c = objective(para_of(cand))
        if loss_c < losses[0]:
            cur, found, loss_best, ad_best = cand, True, loss_c, ad_c
            break
        alpha /= 2
    return cur, loss_best, ad_best, found
