"""Drop-in for the solver-facing part of the reference's RPModule/rputil.py.

Only ``opts`` (rputil.py:11-22) is on the hot path; it keeps the reference's field
names and defaults because callers mutate the fields freely
(evaluation.py:97-106, trainRelativePoseModuleRecFD.py:254-257)."""
import numpy as np


class opts(object):
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


def angular_distance_np(R_hat, R):
    """Rotation angle between two (batches of) rotations in degrees (rputil.py:24-35)."""
    R_hat = np.asarray(R_hat).reshape(-1, 3, 3)
    R = np.asarray(R).reshape(-1, 3, 3)
    tr = np.einsum('nij,nij->n', R_hat, R)
    return np.arccos(((tr - 1) / 2).clip(-1, 1)) / np.pi * 180.0
