"""The FD-trainer objective (trainRelativePoseModuleRecFD.py:215-233) as one batched launch
against the same loop over the numpy oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_objective(prims, sig):
    from oracle import rp_oracle
    from relativepose_b200 import synth
    from relativepose_b200.RPModule.rputil import angular_distance_np
    loss = ad = 0.0
    for p in prims:
        s, t = synth.record_to_dicts(p)
        R_hat = rp_oracle.solve_pair(s, t, rp_oracle.Params(*sig))
        loss += np.power(R_hat[:3, :3] - p['R_gt'][:3, :3], 2).sum()
        ad += angular_distance_np(R_hat[:3, :3].reshape(1, 3, 3), p['R_gt'][:3, :3].reshape(1, 3, 3))[0]
    return loss / len(prims), ad / len(prims)


def test_objective_matches_oracle_loop(tmp_path):
    from relativepose_b200 import fd_objective, synth
    from relativepose_b200.RPModule.rputil import opts
    prims = [synth.make_pair(900 + i, n) for i, n in enumerate((30, 41, 26, 52, 37, 33))]
    path = os.path.join(str(tmp_path), "primitives.npy")
    fd_objective.save_primitives(path, prims)
    loaded = fd_objective.load_primitives(path)
    assert len(loaded) == len(prims) and np.array_equal(loaded[2]['feat_src'], prims[2]['feat_src'])
    obj = fd_objective.Objective(loaded)
    for sig in (synth.shipped_params('suncg')[0], (0.523 / 2, 0.523 / 2, 0.08 / 2, 0.01)):
        sig = tuple(float(x) for x in sig)
        loss, ad = obj(opts(*sig))
        lo, ao = _oracle_objective(prims, sig)
        print("objective: gpu (%.6e, %.6f deg)  oracle (%.6e, %.6f deg)" % (loss, ad, lo, ao))
        assert abs(loss - lo) <= 1e-8 * max(1.0, abs(lo)) and abs(ad - ao) <= 1e-6
    assert obj.evaluations == 2


def test_fd_step_descends_or_keeps():
    from relativepose_b200 import fd_objective, synth
    from relativepose_b200.RPModule.rputil import opts
    prims = [synth.make_pair(700 + i, 40, pos_noise=0.02) for i in range(24)]
    obj = fd_objective.Objective(prims)
    cur = np.array([0.523 / 2, 0.523 / 2, 0.08 / 2, 0.01])
    l0, _ = obj(opts(*cur))
    new, lb, ab, found = fd_objective.fd_step(obj, cur, np.random.RandomState(0))
    assert np.isfinite(new).all() and lb <= l0 + 1e-15
    assert obj.evaluations >= 11          # 1 + 10 probes (+ line search): every one a single launch over the cache
