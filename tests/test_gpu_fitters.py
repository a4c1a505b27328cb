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


def test_near_degenerate_affinity_graph():
    """Two weakly coupled groups of mutually consistent correspondences (two rigid motions of similar support): the two
    leading eigenvalues of the affinity differ by ~0.3 %, where a plain power iteration needs thousands of steps
    (the reference's ARPACK call does not care).  The accelerated iteration must converge within a few dozen steps per
    alternation and reproduce the oracle's (scipy eigs) pose."""
    from oracle import rp_oracle
    from relativepose_b200 import synth
    from relativepose_b200 import solver as S
    from RPModule import rpmodule as M
    rs = np.random.RandomState(3)
    groups, geo, edges, wts = [], [], [], []
    base = 0
    for gi, n in enumerate((40, 36)):
        T = synth.make_pose(20 + gi)
        sp = rs.uniform(-3, 3, (n, 3))
        sn = rs.randn(n, 3); sn /= np.linalg.norm(sn, axis=1, keepdims=True)
        tp = sp @ T[:3, :3].T + T[:3, 3] + rs.randn(n, 3) * 0.003
        tn = sn @ T[:3, :3].T
        geo.append((sp, tp, sn, tn))
        iu = np.array([(a, b) for a in range(n) for b in range(a + 1, n) if rs.rand() < 0.35])
        w = rs.uniform(0.5, 1.0, len(iu))
        Wd = np.zeros((n, n)); Wd[iu[:, 0], iu[:, 1]] = w; Wd = Wd + Wd.T
        groups.append((iu + base, w, np.linalg.eigvalsh(Wd)[-1]))
        base += n
    scale = 0.997 * groups[0][2] / groups[1][2]                  # second group's leading eigenvalue = 99.7 % of the first's
    pairs = np.concatenate((groups[0][0], groups[1][0], np.array([[3, 45], [10, 60], [22, 70]])))
    w = np.concatenate((groups[0][1], groups[1][1] * scale, np.full(3, 1e-3)))
    SP, TP, SN, TN = [np.concatenate([g[k] for g in geo]) for k in range(4)]
    i1, i2 = pairs[:, 0], pairs[:, 1]
    rows = rp_oracle._Rows(np.concatenate((SP[i1], SP[i2])), np.concatenate((TP[i1], TP[i2])),
                           np.concatenate((SN[i1], SN[i2])), np.concatenate((TN[i1], TN[i2])), w)
    w2 = np.concatenate((w, w))
    n_nodes = base
    mine = M.fit_irls_sm(rows.SP, rows.TP, rows.SN, rows.TN, w2, w2.copy(), w, 0.3, i1, i2, n_nodes, 1)
    stats = S.default_solver().last_fit_stats[0]
    theirs = rp_oracle.fit_irls_sm(rows, 0.3, i1, i2, n_nodes)
    print("near-degenerate graph: eigen iterations total %d, max per alternation %d, hit cap %d; |T - T_oracle| = %.2e"
          % (stats[4], stats[5], stats[6], np.linalg.norm(mine - theirs)))
    assert stats[6] == 0 and stats[5] <= 400, "eigen iteration did not converge quickly"
    assert np.linalg.norm(mine - theirs) <= 1e-7
