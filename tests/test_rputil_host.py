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


def test_matching_primitive_needs_keypoints():
    import pytest
    from RPModule.rpmodule import getMatchingPrimitive
    with pytest.raises(NotImplementedError):
        getMatchingPrimitive({}, {}, "suncg", "skybox", True)
