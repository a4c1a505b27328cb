"""rputil.getPixel (host numpy) against the reference's own function (tests/golden/make_rputil_golden.py)."""
import os

import numpy as np


def test_getpixel_matches_reference():
    from RPModule.rputil import getPixel
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rputil_golden.npz"))
    for ds in ("suncg", "matterport"):
        pc, nn = getPixel(G["depth"], G["normal"], G["pts"], dataset=ds)
        assert pc.shape == (3, 60) and nn.shape == (60, 3)
        assert np.abs(pc - G["pc_" + ds]).max() <= 1e-12 and np.abs(nn - G["nn_" + ds]).max() <= 1e-12


def test_matching_primitive_default_keypoint_stage_is_the_reference_one():
    """Without keypoint_fn the reference's own stage runs (rputil.getKeypoint: it reads dataS['rgb'] first); an unknown
    dataset is rejected before any work."""
    import pytest
    from RPModule.rpmodule import getMatchingPrimitive
    with pytest.raises(KeyError):
        getMatchingPrimitive({}, {}, "suncg", "skybox", True)
    with pytest.raises(ValueError):
        getMatchingPrimitive({}, {}, "nyu", "skybox", True)
