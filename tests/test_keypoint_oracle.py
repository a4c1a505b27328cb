"""oracle/keypoint_oracle.py (restatement of rputil.getKeypoint / getKeypoint_kinect / Sampling, rputil.py:141-371) against
the golden outputs of the unmodified reference (tests/golden/make_keypoint_golden.py): bit exact."""
import os

import numpy as np
import pytest

from oracle import keypoint_oracle as ko
from relativepose_b200 import synth

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keypoint_golden.npz"))
NAMES = [str(n) for n in G['names']]
KEYS = ('pts', 'ptsNorm', 'ptsW', 'ptt', 'pttNorm', 'pttW')


def case(name):
    kinect, seed = [int(v) for v in G[name + '/meta']]
    return kinect, seed, synth.make_feature_map(seed + 200), synth.make_feature_map(seed + 300)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    kinect, seed, fs, ft = case(name)
    out = ko.get_keypoint(ko.place_sift(G[name + '/kps'], kinect), ko.place_sift(G[name + '/kpt'], kinect), fs, ft,
                          np.random.RandomState(seed + 400), bool(kinect))
    for k, o in zip(KEYS, out):
        assert np.array_equal(o, G[name + '/' + k]), k


def test_sampling_suppression_window_semantics():
    d = np.full((1, 40, 50), 5.0, np.float32)
    d[0, 10, 12] = 0.1          # best
    d[0, 12, 20] = 0.2          # inside the 15-pixel window of the best -> suppressed
    d[0, 30, 45] = 0.3          # outside
    pt = ko.sampling(d, 2)
    assert pt[0, 0].tolist() == [12, 10] and pt[0, 1].tolist() == [45, 30]
