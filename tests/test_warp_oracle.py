"""oracle/warp_oracle.py (numpy restatement of util.warping / reproj_helper / Pano2PointCloud, util.py:94-172, 537-811)
against the golden vectors the unmodified reference produced (tests/golden/make_warp_golden.py) and, when the reference
tree is present, against the live reference bit for bit."""
import os

import numpy as np
import pytest

from oracle import ref_loader, warp_oracle
from relativepose_b200 import synth

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_golden.npz"))
NAMES = [str(n) for n in G['names']]


def case_inputs(name):
    seed, inv = [int(v) for v in G[name + '/meta']]
    ds = str(G[name + '/dataset'])
    view = synth.make_warp_view(seed, ds)
    R = synth.make_pose(seed)
    if inv:
        R = np.linalg.inv(R)
    return view, R, ds


def check_against_golden(name, out):
    """out [8, 102400] (float32 or float64).  Exact on the validity mask; values to float32 rounding."""
    m = out[7] != 0
    assert np.array_equal(np.packbits(m), G[name + '/mask']), "validity mask differs from the reference"
    idx = G[name + '/idx']
    assert np.array_equal(out[:, idx].astype(np.float32), G[name + '/vals'])
    assert np.allclose(out.astype(np.float64).sum(1), G[name + '/sums'], rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    view, R, ds = case_inputs(name)
    out = warp_oracle.warping(view, R, ds)
    assert out.shape == (1, 8, 160, 640) and out.dtype == np.float64
    check_against_golden(name, out[0].reshape(8, -1))
    assert np.array_equal(out[0].reshape(8, -1).sum(1), G[name + '/sums'])        # same machine arithmetic: bit exact


def test_identity_pose_gives_zeros():
    view = synth.make_warp_view(3, 'suncg')
    assert not warp_oracle.warping(view, np.eye(4), 'suncg').any()


@pytest.mark.parametrize("ds", ['suncg', 'matterport', 'scannet'])
def test_pano2pointcloud_golden(ds):
    full = np.random.RandomState(9).uniform(0.5, 5, (160, 640)).astype(np.float32)
    full[np.random.RandomState(10).rand(160, 640) < 0.05] = 0
    pc = warp_oracle.pano2pointcloud(full, ds)
    assert list(pc.shape) == list(G['pano_' + ds + '/shape'])
    assert np.array_equal(pc[:, ::97], G['pano_' + ds + '/sample'])
    assert np.allclose(pc.sum(1), G['pano_' + ds + '/sums'], rtol=1e-12)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("ds", ['suncg', 'matterport', 'scannet'])
def test_oracle_matches_live_reference(ds):
    util = ref_loader.load_reference_rpmodule().util
    for seed in (7, 8):
        view, R = synth.make_warp_view(seed, ds), synth.make_pose(seed)
        assert np.array_equal(util.warping(view, R, ds), warp_oracle.warping(view, R, ds))
    d = synth.make_warp_view(5, ds)[0, 6]
    assert np.array_equal(util.Pano2PointCloud(d, ds), warp_oracle.pano2pointcloud(d, ds))
