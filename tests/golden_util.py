"""Helpers to read tests/golden/rp_golden.npz (made by tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rp_golden.npz")
_IN = ('pc_src', 'normal_src', 'feat_src', 'weight_src', 'pc_tgt', 'normal_tgt', 'feat_tgt', 'weight_tgt', 'R_gt')


class Case(object):
    def __init__(self, blob, name):
        self.name = name
        self.record = {k: blob[name + '/in/' + k] for k in _IN}
        m = blob[name + '/meta']
        self.sigmas = m[:4]
        self.topk = int(m[4])
        self.method = str(blob[name + '/method'])
        self.T = blob[name + '/out/T']
        self.topk_sets = blob[name + '/out/topk'] if name + '/out/topk' in blob else None
        self.row = blob[name + '/out/row'] if name + '/out/row' in blob else None
        self.col = blob[name + '/out/col'] if name + '/out/col' in blob else None
        self.w = blob[name + '/out/w'] if name + '/out/w' in blob else None

    def dicts(self):
        r = self.record
        s = {'pc': r['pc_src'], 'normal': r['normal_src'], 'feat': r['feat_src'], 'weight': r['weight_src']}
        t = {'pc': r['pc_tgt'], 'normal': r['normal_tgt'], 'feat': r['feat_tgt'], 'weight': r['weight_tgt']}
        return s, t

    def apply(self, para):
        para.sigmaAngle1, para.sigmaAngle2, para.sigmaDist, para.sigmaFeat = [float(x) for x in self.sigmas]
        para.topK = self.topk
        para.method = self.method
        return para

    def pair_keys(self):
        """Surviving pairs as sorted array of (row,col) flat-id tuples (rpmodule.py:495-496)."""
        if self.row is None:
            return None
        return np.stack((self.row, self.col), 1)


def load_cases():
    blob = np.load(GOLDEN, allow_pickle=False)
    return [Case(blob, str(n)) for n in blob['names']]


def case_names():
    blob = np.load(GOLDEN, allow_pickle=False)
    return [str(n) for n in blob['names']]
