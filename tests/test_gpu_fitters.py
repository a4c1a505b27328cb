"""Module-level fitters (rpmodule.py:17,60,86,169,212) through the rp_spectral_irls_solve stage entry, against the
oracle's restatement of the same functions on the stacked rows the helper builds (rpmodule.py:474-489)."""
import numpy as np
import pytest

from tests.golden_util import load_cases

pytestmark = pytest.mark.gpu
CASES = {c.name: c for c in load_cases()}


def _stacked(case):
    from oracle import rp_oracle
    s, t = case.dicts()
    tr = {}
    rp_oracle.solve_pair(s, t, case.apply(rp_oracle.Params()), tr)
    cor, pairs, w = tr['corres'], tr['pairs'], tr['w']
    i1, j1 = cor[0, pairs[:, 0]], cor[1, pairs[:, 0]]
    i2, j2 = cor[0, pairs[:, 1]], cor[1, pairs[:, 1]]
    Ps, Pt, Ns, Nt = s['pc'], t['pc'], s['normal'], t['normal']
    rows = rp_oracle._Rows(np.concatenate((Ps[i1], Ps[i2])), np.concatenate((Pt[j1], Pt[j2])),
                           np.concatenate((Ns[i1], Ns[i2])), np.concatenate((Nt[j1], Nt[j2])), w)
    n_t = Pt.shape[0]
    return rows, i1 * n_t + j1, i2 * n_t + j2, Ps.shape[0], n_t


@pytest.mark.parametrize("name", ["n52_s0", "n103_s1", "rag_40_70"])
def test_fitters_match_oracle(name):
    from oracle import rp_oracle
    from RPModule import rpmodule as M
    rows, row, col, ns, nt = _stacked(CASES[name])
    w2 = np.concatenate((rows.w_pair, rows.w_pair))
    a = (rows.SP, rows.TP, rows.SN, rows.TN, w2, w2.copy())
    mu = 0.3
    for mine, theirs in ((M.fit_horn87(*a, mu), rp_oracle.fit_horn87(rows, mu)),
                         (M.fit_irls(*a, mu), rp_oracle.fit_irls(rows, mu)),
                         (M.fit_spectral(*a, rows.w_pair, mu, row, col, ns, nt), rp_oracle.fit_spectral(rows, mu, row, col, ns * nt)),
                         (M.fit_irls_sm(*a, rows.w_pair, mu, row, col, ns, nt), rp_oracle.fit_irls_sm(rows, mu, row, col, ns * nt))):
        assert mine.shape == (4, 4)
        assert np.linalg.norm(mine - theirs) <= 1e-8, np.linalg.norm(mine - theirs)


def test_horn87_np_matches_oracle():
    from oracle import rp_oracle
    from RPModule.rpmodule import horn87_np
    rs = np.random.RandomState(0)
    src, tgt, w = rs.randn(3, 3, 50), rs.randn(3, 3, 50), rs.rand(3, 50)
    R = horn87_np(src, tgt, w)
    assert R.shape == (3, 3, 3)
    for k in range(3):
        assert np.abs(R[k] - rp_oracle.horn_rotation(src[k], tgt[k], w[k])).max() <= 1e-9
    R1 = horn87_np(src[0], tgt[0])
    assert np.abs(R1[0] - rp_oracle.horn_rotation(src[0], tgt[0], np.ones(50))).max() <= 1e-9
