"""GPU parity of the view warp (csrc/rp_warp.cu) with the reference's util.warping:
against the committed goldens of the unmodified reference and against the numpy oracle on the full image.
Bar: the validity mask (which target pixels receive a point) and the winner of every collision are exact; values equal
the reference's float64 results after its caller's float32 cast (bit exact, asserted with max-abs 0 on rgb and <= 1 ulp
of float32 elsewhere)."""
import types

import numpy as np
import pytest

from tests.test_warp_oracle import NAMES, case_inputs, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", NAMES)
def test_warp_matches_reference(name):
    import torch
    from oracle import warp_oracle
    from relativepose_b200 import util
    view, R, ds = case_inputs(name)
    out = util.warping_device(torch.from_numpy(view).cuda(), R[None], ds)[0].cpu().numpy().reshape(8, -1)
    check_against_golden(name, out)
    ref = warp_oracle.warping(view, R, ds)[0].reshape(8, -1)
    assert np.array_equal(out[7] != 0, ref[7] != 0)
    assert np.array_equal(out[0:3], ref[0:3].astype(np.float32))                   # colours are copied: winner identity
    d = np.abs(out.astype(np.float64) - ref)
    assert d.max() <= 1e-6, d.max()
    exact = float((out == ref.astype(np.float32)).mean())
    print("%s: %d pixels written, %.6f of all values bit-equal to float32(reference)" % (name, int((ref[7] != 0).sum()), exact))
    assert exact >= 0.99999


def test_warp_batch_identity_and_numpy_surface():
    import torch
    import util as root_util                       # repo-root shim named like the reference's module
    from oracle import warp_oracle
    from relativepose_b200 import synth
    views = np.concatenate([synth.make_warp_view(s, 'matterport') for s in (0, 1, 2)])
    Rs = np.stack([synth.make_pose(0), np.eye(4), np.linalg.inv(synth.make_pose(2))])
    out = root_util.warping_device(torch.from_numpy(views).cuda(), Rs, 'matterport').cpu().numpy()
    assert not out[1].any()                                                         # util.py:95-96
    for b in (0, 2):
        ref = warp_oracle.warping(views[b:b + 1], Rs[b], 'matterport')[0]
        assert np.array_equal(out[b, 7] != 0, ref[7] != 0) and np.abs(out[b] - ref).max() <= 1e-6
    one = root_util.warping(views[0:1], Rs[0], 'matterport')
    assert one.dtype == np.float64 and one.shape == (1, 8, 160, 640) and np.array_equal(one[0].astype(np.float32), out[0])


@pytest.mark.parametrize("ds", ['suncg', 'matterport', 'scannet'])
def test_pano2pointcloud(ds):
    from oracle import warp_oracle
    from relativepose_b200 import util
    full = np.random.RandomState(9).uniform(0.5, 5, (160, 640)).astype(np.float32)
    full[np.random.RandomState(10).rand(160, 640) < 0.05] = 0
    pc = util.Pano2PointCloud(full, ds)
    ref = warp_oracle.pano2pointcloud(full, ds)
    assert pc.shape == ref.shape and np.array_equal(pc, ref)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_blend_completion(dt):
    import torch
    from oracle import warp_oracle
    from relativepose_b200 import util
    rs = np.random.RandomState(4)
    f = rs.randn(2, 54, 160, 640).astype(np.float32)
    mask = np.zeros((2, 160, 640), np.float32); mask[:, :, 160:320] = 1
    ng = rs.randn(2, 160, 640, 3); ng /= np.linalg.norm(ng, axis=3, keepdims=True)
    ng, dg = ng.astype(dt), rs.uniform(0.5, 5, (2, 160, 640)).astype(dt)
    nrm, dep = util.blend_completion_device(torch.from_numpy(f).cuda(), torch.from_numpy(mask), torch.from_numpy(ng), torch.from_numpy(dg))
    assert nrm.dtype == (torch.float64 if dt == np.float64 else torch.float32)
    for b in range(2):
        rn, rd = warp_oracle.blend_completion(f[b], mask[b][:, :, None], ng[b], dg[b])
        tol = 1e-15 if dt == np.float64 else 2e-7
        assert np.abs(nrm[b].cpu().numpy() - rn).max() <= tol and np.array_equal(dep[b].cpu().numpy(), rd)


def test_via_completion_two_steps_gpu_warp_equals_oracle_warp():
    """RelativePoseEstimationViaCompletion with alterStep=2: the second step warps both scans with the first estimate.
    Running it with the GPU warp and with the numpy oracle's warp injected must give the same pose."""
    import torch
    from oracle import warp_oracle
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    from RPModule.rpmodule import RelativePoseEstimationViaCompletion
    from RPModule.rputil import opts
    a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
    torch.manual_seed(0)
    net = SCNet(a).cuda()
    rs = np.random.RandomState(11)

    def scan(seed):
        v = synth.make_warp_view(seed, 'suncg')
        full = np.random.RandomState(seed).rand(160, 640)
        yy, xx = np.mgrid[0:160, 0:640]
        depth = 2.5 + 1.5 * np.sin(xx / 37.0 + seed) * np.cos(yy / 23.0) + 0.2 * full
        nrm = rs.randn(160, 640, 3); nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
        return {'rgb': rs.uniform(0, 1, (160, 640, 3)), 'norm': nrm, 'depth': depth}

    def keypoints(dataS, dataT, dataset):
        r2 = np.random.RandomState(5)

        def grid(n):
            p = np.stack((r2.uniform(1, 637, n), r2.uniform(1, 157, n)), 1)
            return p, p / np.array([640.0, 160.0]), np.where((p[:, 0] >= 160) & (p[:, 0] <= 320), 1.0, 0.99)
        return grid(40) + grid(45)

    P = synth.shipped_params('suncg')
    para = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
    args = types.SimpleNamespace(snumclass=15, featureDim=32, outputType='rgbdnsf', maskMethod='second', alterStep=2,
                                 dataset='suncg', para=para, representation='skybox', completion=True)
    s, t = scan(1), scan(2)
    calls = []

    def oracle_warp(view, R, dataset):
        calls.append(1)
        return warp_oracle.warping(view, R, dataset)
    T_gpu = RelativePoseEstimationViaCompletion(net, s, t, args, keypoint_fn=keypoints)
    T_orc = RelativePoseEstimationViaCompletion(net, s, t, args, keypoint_fn=keypoints, warping_fn=oracle_warp)
    print("two-step completion: |T_gpu - T_oracle_warp| = %.3e, oracle warps used: %d" % (np.linalg.norm(T_gpu - T_orc), len(calls)))
    assert np.isfinite(T_gpu).all() and np.linalg.norm(T_gpu - T_orc) <= 1e-6


def test_warp_batch_invariance_full_size():
    """64 views in one call (configs[3]: 32 ScanNet pairs) == the same views one by one, bit for bit (the winner map is per
    view; nothing crosses views)."""
    import torch
    from relativepose_b200 import synth, util
    views = torch.from_numpy(np.concatenate([synth.make_warp_view(s % 5, 'scannet') for s in range(64)])).cuda()
    Rs = np.stack([synth.make_pose(s) if s % 7 else np.eye(4) for s in range(64)])
    big = util.warping_device(views, Rs, 'scannet')
    for b in (0, 1, 7, 33, 63):
        one = util.warping_device(views[b:b + 1], Rs[b:b + 1], 'scannet')
        assert torch.equal(big[b:b + 1], one)
    assert not big[0].any() and not big[7].any() and big[1].any()          # identity poses -> zeros (util.py:95-96)
