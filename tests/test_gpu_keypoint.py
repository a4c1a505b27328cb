"""GPU parity of the keypoint augmentation (csrc/rp_keypoint.cu, rputil.getKeypoint / getKeypoint_kinect / Sampling;
SURVEY.md section 8f row 2) with the unmodified reference's outputs (goldens) and the numpy oracle.  The SIFT detections
recorded with the goldens are injected (the detector is OpenCV on the CPU, outside the GPU path); random draws come from a
RandomState seeded like the golden run.  Bar: every keypoint coordinate and weight identical."""
import numpy as np
import pytest

from tests.test_keypoint_oracle import G, KEYS, NAMES, case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", NAMES)
def test_get_keypoint_matches_reference(name):
    import torch
    from RPModule import rputil
    kinect, seed, fs, ft = case(name)
    seq = [G[name + '/kps'], G[name + '/kpt']]
    calls = []

    def sift_fn(gray):                       # hand back the recorded detections in call order (source, target)
        calls.append(gray.shape)
        return seq[len(calls) - 1]
    rng = np.random.RandomState(seed + 400)
    rgb = np.zeros((160, 640, 3), np.uint8)
    full = np.zeros((480, 640, 3), np.uint8)
    fsd, ftd = torch.from_numpy(fs).cuda(), torch.from_numpy(ft).cuda()
    if kinect:
        out = rputil.getKeypoint_kinect(rgb, rgb, fsd, ftd, full, full, rng=rng, sift_fn=sift_fn)
        assert calls == [(480, 640), (480, 640)]
    else:
        out = rputil.getKeypoint(rgb, rgb, fsd, ftd, rng=rng, sift_fn=sift_fn)
        assert calls == [(160, 160), (160, 160)]
    for k, o in zip(KEYS, out):
        g = G[name + '/' + k]
        assert o.shape == g.shape, k
        assert np.array_equal(o, g), "%s: %d of %d entries differ" % (k, int((o != g).sum()), g.size)


def test_sampling_and_fused_match_sample_agree_with_oracle():
    import torch
    from oracle import keypoint_oracle as ko
    from RPModule import rputil
    from relativepose_b200 import synth
    feat = synth.make_feature_map(77)
    rs = np.random.RandomState(3)
    ptn = np.stack((rs.uniform(0.01, 0.98, 37), rs.uniform(0.01, 0.98, 37)), 1)
    q = ko.interpolate(synth.make_feature_map(78), ptn)                                  # [32, 37]
    ref = ko.sampling(ko.dense_dist(q, feat), 3)
    got = rputil.match_sample(torch.from_numpy(q).cuda(), torch.from_numpy(feat).cuda(), 3)
    assert np.array_equal(got, ref)
    d = ko.dense_dist(q[:, :9], feat)
    assert np.array_equal(rputil.Sampling(d, 2), ko.sampling(d, 2))
    assert np.array_equal(rputil.Sampling(torch.from_numpy(d).cuda(), 2), ko.sampling(d, 2))


def test_get_keypoint_no_detections():
    import torch
    from RPModule import rputil
    f = torch.zeros((32, 160, 640), device='cuda')
    out = rputil.getKeypoint(np.zeros((160, 640, 3), np.uint8), np.zeros((160, 640, 3), np.uint8), f, f, sift_fn=lambda g: np.zeros((0, 2)))
    assert out == (None,) * 6


def test_match_sample_is_independent_of_the_query_batch():
    """100 queries in one call == the same queries in chunks (queries never interact; slices are combined in index order)."""
    import torch
    from RPModule import rputil
    from relativepose_b200 import synth
    feat = torch.from_numpy(synth.make_feature_map(5)).cuda()
    q = torch.tanh(torch.randn((32, 100), generator=torch.Generator().manual_seed(1))).cuda()
    full = rputil.match_sample(q, feat, 2)
    parts = np.concatenate([rputil.match_sample(q[:, i:i + 17].contiguous(), feat, 2) for i in range(0, 100, 17)])
    assert full.shape == (100, 2, 2) and np.array_equal(full, parts)
    assert (full[:, 0] != full[:, 1]).any(axis=1).all()                      # the second pick lies outside the first one's window
