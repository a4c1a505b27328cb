"""The device-resident batched pipeline (relativepose_b200/pipeline.py; BASELINE configs[2]/[3]) against the single-pair
reference-named path: the gather kernel vs rputil.getPixel + rputil.interpolate, solve_from_maps vs RelativePoseEstimation,
and the batched three-stage alternation vs RelativePoseEstimationViaCompletion pair by pair."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _keypoints(seed, n_img, K):
    rs = np.random.RandomState(seed)
    pts = np.stack((rs.uniform(1, 637, (n_img, K)), rs.uniform(1, 157, (n_img, K))), 2)
    w = np.where((pts[..., 0] >= 160) & (pts[..., 0] <= 320), 1.0, 0.99)
    return pts, w


def _maps(seed, n_img):
    rs = np.random.RandomState(seed)
    nrm = rs.randn(n_img, 160, 640, 3)
    nrm /= np.linalg.norm(nrm, axis=3, keepdims=True)
    return rs.uniform(0.5, 5, (n_img, 160, 640)), nrm


@pytest.mark.parametrize("dataset", ["suncg", "matterport"])
def test_gather_primitives_matches_getpixel_and_interpolate(dataset):
    import torch
    from relativepose_b200 import pipeline, synth
    from RPModule.rputil import getPixel, interpolate
    n_img, K = 4, 57
    pts, w = _keypoints(1, n_img, K)
    depth, nrm = _maps(2, n_img)
    feat = torch.from_numpy(np.stack([synth.make_feature_map(10 + i) for i in range(n_img)])).cuda()
    big = torch.zeros((n_img, 54, 160, 640), device='cuda')
    big[:, 22:54] = feat                                                   # a channel slice of a wider tensor, like the net output
    d = pipeline.gather_primitives(big[:, 22:54], torch.from_numpy(depth), torch.from_numpy(nrm), pts, w, dataset)
    for b in range(2):
        for side, (pc_d, nn_d, ft_d, w_d) in enumerate(((d.pc_s, d.nrm_s, d.feat_s, d.w_s), (d.pc_t, d.nrm_t, d.feat_t, d.w_t))):
            i = 2 * b + side
            pc, nn = getPixel(depth[i], nrm[i], pts[i], dataset=dataset)
            des = interpolate(feat[i], pts[i] / np.array([640.0, 160.0])).cpu().numpy().T
            sl = slice(b * K, (b + 1) * K)
            assert np.abs(pc_d[sl].cpu().numpy() - pc.T).max() <= 1e-12
            assert np.abs(nn_d[sl].cpu().numpy() - nn).max() <= 1e-12
            assert np.array_equal(ft_d[sl].cpu().numpy(), des)
            assert np.array_equal(w_d[sl].cpu().numpy(), w[i])


def test_solve_from_maps_equals_single_pair_calls():
    import torch
    from relativepose_b200 import pipeline, synth
    from RPModule.rpmodule import RelativePoseEstimation
    from RPModule.rputil import opts
    B, K = 3, 60
    pts, w = _keypoints(3, 2 * B, K)
    depth, nrm = _maps(4, 2 * B)
    feat = torch.from_numpy(np.stack([synth.make_feature_map(30 + i) for i in range(2 * B)])).cuda()
    para = opts(*synth.shipped_params('suncg')[0])
    para.sigmaFeat = 0.05
    T = pipeline.solve_from_maps(feat, torch.from_numpy(depth), torch.from_numpy(nrm), pts, w, para, 'suncg')
    for b in range(B):
        def kp(dS, dT, ds, b=b):
            return (pts[2 * b], pts[2 * b] / np.array([640.0, 160.0]), w[2 * b], pts[2 * b + 1], pts[2 * b + 1] / np.array([640.0, 160.0]), w[2 * b + 1])
        data = [{'rgb': None, 'depth': depth[2 * b + s], 'normal': nrm[2 * b + s], 'feat': feat[2 * b + s]} for s in (0, 1)]
        T1 = RelativePoseEstimation(data[0], data[1], para, 'suncg', 'skybox', 'second', keypoint_fn=kp)
        assert np.linalg.norm(T[b] - T1) <= 1e-8, (b, np.linalg.norm(T[b] - T1))


def test_via_completion_batch_equals_pairwise():
    import torch
    from relativepose_b200 import pipeline, synth
    from relativepose_b200.model.mymodel import SCNet
    from RPModule.rpmodule import RelativePoseEstimationViaCompletion
    from RPModule.rputil import opts
    a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
    torch.manual_seed(0)
    net = SCNet(a).cuda()
    B, K = 2, 45
    rs = np.random.RandomState(5)
    yy, xx = np.mgrid[0:160, 0:640]
    rgb = rs.uniform(0, 1, (2 * B, 160, 640, 3))
    depth, nrm = _maps(6, 2 * B)
    depth = np.stack([2.5 + 1.5 * np.sin(xx / 37.0 + i) * np.cos(yy / 23.0) + 0.2 * depth[i] / 5 for i in range(2 * B)])
    pts, w = _keypoints(7, 2 * B, K)
    P = synth.shipped_params('suncg')
    para = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
    args = types.SimpleNamespace(snumclass=15, featureDim=32, outputType='rgbdnsf', maskMethod='second', alterStep=2,
                                 dataset='suncg', para=para, representation='skybox', completion=True)
    T = pipeline.RelativePoseEstimationViaCompletion_batch(net, rgb, nrm, depth, pts, w, args)
    assert T.shape == (B, 4, 4)
    # network one pair at a time (scans uploaded chunk by chunk on the copy stream), one solve per step over both pairs: the same bits
    Tc = pipeline.RelativePoseEstimationViaCompletion_batch(net, rgb, nrm, depth, pts, w, args, chunk=1)
    assert np.array_equal(T, Tc)
    for b in range(B):
        def kp(dS, dT, ds, b=b):
            return (pts[2 * b], pts[2 * b] / np.array([640.0, 160.0]), w[2 * b], pts[2 * b + 1], pts[2 * b + 1] / np.array([640.0, 160.0]), w[2 * b + 1])
        ds_ = [{'rgb': rgb[2 * b + s], 'norm': nrm[2 * b + s], 'depth': depth[2 * b + s]} for s in (0, 1)]
        T1 = RelativePoseEstimationViaCompletion(net, ds_[0], ds_[1], args, keypoint_fn=kp)
        print("pair %d: |T_batch - T_single| = %.3e" % (b, np.linalg.norm(T[b] - T1)))
        assert np.linalg.norm(T[b] - T1) <= 1e-6
