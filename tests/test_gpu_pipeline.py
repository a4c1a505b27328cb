"""Net -> descriptors -> solver on the GPU (BASELINE configs[2]/[3] shape): SCNet forward, rputil.interpolate at fixed
keypoints, rputil.getPixel, RelativePoseEstimation; the oracle solver runs on the SAME primitives, so the comparison
isolates the solver from the (separately tested) network tolerance."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_interpolate_matches_reference():
    import torch
    from RPModule.rputil import interpolate
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rputil_golden.npz"))
    rs = np.random.RandomState(int(G["feat_seed"]))
    feat = rs.randn(32, 160, 640).astype(np.float32)
    out = interpolate(torch.from_numpy(feat).cuda(), G["ptn"]).cpu().numpy()
    assert out.shape == (32, 77)
    assert np.abs(out - G["interp"]).max() <= 1e-6


def test_scnet_to_solver_pipeline():
    import torch
    from oracle import rp_oracle
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    from RPModule.rpmodule import RelativePoseEstimation
    from RPModule.rputil import opts
    a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
    torch.manual_seed(0)
    net = SCNet(a).cuda()
    x = torch.from_numpy(synth.make_panorama_pair(21, "suncg")).cuda()
    f = net(x)                                             # [2,54,160,640]
    i0 = 3 + 3 + 1 + 15                                    # evaluation.py:137-138
    rs = np.random.RandomState(3)
    captured = {}

    def keypoints(dataS, dataT, dataset):                  # fixed jittered grid instead of SIFT
        def grid(n):
            p = np.stack((rs.uniform(1, 637, n), rs.uniform(1, 157, n)), 1)
            return p, p / np.array([640.0, 160.0]), np.where((p[:, 0] >= 160) & (p[:, 0] <= 320), 1.0, 0.99)
        ps, psn, pw = grid(60)
        pt, ptn, tw = grid(55)
        return ps, psn, pw, pt, ptn, tw

    def data(k):
        depth = np.abs(f[k, 6].cpu().numpy().astype(np.float64)) + 0.5
        nrm = f[k, 3:6].cpu().numpy().astype(np.float64).transpose(1, 2, 0)
        nrm /= (np.linalg.norm(nrm, axis=2, keepdims=True) + 1e-12)
        return {'rgb': None, 'depth': depth, 'normal': nrm, 'feat': f[k, i0:i0 + 32]}

    P = synth.shipped_params('suncg')
    para = opts(*P[0])
    para.sigmaFeat = 0.05          # untrained descriptors: widen the soft-match kernel so rows do not all underflow
    import relativepose_b200.RPModule.rpmodule as M
    orig = M.RelativePoseEstimation_helper

    def spy(s, t, p):
        captured['s'], captured['t'] = s, t
        return orig(s, t, p)
    M.RelativePoseEstimation_helper = spy
    try:
        T = RelativePoseEstimation(data(0), data(1), para, 'suncg', 'skybox', 'second', keypoint_fn=keypoints)
    finally:
        M.RelativePoseEstimation_helper = orig
    assert T.shape == (4, 4)
    op = rp_oracle.Params(*P[0]); op.sigmaFeat = 0.05
    tr = {}
    To = rp_oracle.solve_pair(captured['s'], captured['t'], op, tr)
    print("pipeline: status", tr['status'], "pairs", tr.get('n_angle'), "|T-To|", np.linalg.norm(T - To))
    assert np.linalg.norm(T - To) <= 1e-8       # needs the sequential float32 summation order (transposed 'feat' views)


def test_via_completion_one_step():
    """RelativePoseEstimationViaCompletion (rpmodule.py:569-662) with alterStep=1 (no warping needed) equals the manual
    composition apply_mask -> SCNet -> blend -> RelativePoseEstimation."""
    import torch
    from relativepose_b200.model.mymodel import SCNet
    from RPModule.rpmodule import RelativePoseEstimationViaCompletion
    from RPModule.rputil import opts
    from relativepose_b200 import synth
    a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
    torch.manual_seed(0)
    net = SCNet(a).cuda()
    rs = np.random.RandomState(11)

    def scan():
        nrm = rs.randn(160, 640, 3); nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
        return {'rgb': rs.uniform(0, 1, (160, 640, 3)), 'norm': nrm, 'depth': rs.uniform(0.5, 5, (160, 640))}

    def keypoints(dataS, dataT, dataset):
        r2 = np.random.RandomState(5)
        def grid(n):
            p = np.stack((r2.uniform(1, 637, n), r2.uniform(1, 157, n)), 1)
            return p, p / np.array([640.0, 160.0]), np.where((p[:, 0] >= 160) & (p[:, 0] <= 320), 1.0, 0.99)
        return grid(40) + grid(45)

    P = synth.shipped_params('suncg')
    para = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
    args = types.SimpleNamespace(snumclass=15, featureDim=32, outputType='rgbdnsf', maskMethod='second', alterStep=1,
                                 dataset='suncg', para=para, representation='skybox', completion=True)
    T = RelativePoseEstimationViaCompletion(net, scan(), scan(), args, keypoint_fn=keypoints)
    assert T.shape == (4, 4) and np.isfinite(T).all()
    assert np.allclose(T[3], [0, 0, 0, 1]) and abs(np.linalg.det(T[:3, :3]) - 1) < 1e-6 or np.array_equal(T, np.eye(4))
