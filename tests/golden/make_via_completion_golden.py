"""Golden vectors for RelativePoseEstimationViaCompletion (RPModule/rpmodule.py:569-662) from the UNMODIFIED reference,
executed on CPU in this container:

  * the reference solver module through oracle/ref_loader.py (in-memory syntax shim only), the reference ``SCNet`` class
    (model/mymodel.py) loaded with the ``state_dict`` of this repo's container built under ``torch.manual_seed(0)`` (same
    keys/shapes -- asserted), float32 on the CPU, batch statistics (the reference never calls .eval());
  * ``torch_op.v`` keeps tensors on the CPU and ``Tensor.cuda`` is the identity (the reference hard-codes .cuda());
    ``cv2.xfeatures2d.SIFT_create`` is aliased to ``cv2.SIFT_create`` (contrib namespace is gone in OpenCV 4.13);
  * ``numpy.random.seed`` pins the keypoint augmentation; the SIFT detections of every step are recorded (they are an
    *input* of the repo's keypoint stage: ``sift_fn=``), because OpenCV's detector is outside the parity perimeter.

Scenes: relativepose_b200.synth.make_room_scan_pair (two skybox scans of one textured box room), alterStep = 3, sigmas =
rows 0-2 of the shipped final_param_suncg_rlevel_3.txt (evaluation.py:95-101).  Stored per scene and step: R_hat, the
primitives handed to RelativePoseEstimation_helper (keypoint pixels, 3-D points, normals, descriptors, weights), the
oracle's trace of that solve (top-k sets, surviving-pair counts), a strided sample + per-channel moments of the network
output, and the same sample from the reference module run in float64 (the yardstick for float32 summation-order noise).  Usage: python tests/golden/make_via_completion_golden.py
"""
import importlib.util
import os
import sys
import types

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import rp_oracle  # noqa: E402
from oracle.ref_loader import load_reference_rpmodule  # noqa: E402
from relativepose_b200 import synth  # noqa: E402
from relativepose_b200.model.mymodel import SCNet  # noqa: E402

SCENES = (("room_a", 3, 10, 0.05), ("room_b", 11, 48, 0.05))   # name, scene seed, wall texture tiles (SIFT density), sigmaFeat
ALTER = 3


def main():
    torch.set_num_threads(16)
    ref = load_reference_rpmodule()
    ru = ref._rputil
    cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    to_cpu = lambda var, cuda=True, volatile=False: (torch.from_numpy(var).float() if isinstance(var, np.ndarray) else var.float())  # noqa: E731
    ref.torch_op.v = to_cpu
    ru.torch_op.v = to_cpu
    ref.util.torch_op.v = to_cpu
    torch.Tensor.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location("ref_mymodel", "/root/reference/model/mymodel.py")
    refmodel = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refmodel)

    a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in SCNet(a).state_dict().items()}
    net = refmodel.SCNet(a)
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.load_state_dict(sd)

    # recorders: SIFT detections, network outputs, primitives + poses of every alternation step
    rec = {}
    real_sift = cv2.SIFT_create

    class SiftSpy(object):
        def __init__(self, *a, **k):
            self.s = real_sift(*a, **k)

        def detectAndCompute(self, gray, m):
            kps, des = self.s.detectAndCompute(gray, m)
            rec.setdefault('sift', []).append(np.array([k.pt for k in kps], dtype=np.float64).reshape(-1, 2))
            return kps, des
    cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=SiftSpy)
    orig_helper = ref.RelativePoseEstimation_helper
    orig_gk = ref.getKeypoint

    def gk_spy(*args, **kw):
        rec.setdefault('rng', []).append(np.random.get_state())          # state on entry of every keypoint stage
        out = orig_gk(*args, **kw)
        rec.setdefault('kp', []).append(out)
        return out

    def helper_spy(dataS, dataT, para):
        T = orig_helper(dataS, dataT, para)
        rec.setdefault('prim', []).append((dataS, dataT, para, T))
        return T
    ref.RelativePoseEstimation_helper = helper_spy
    ref.getKeypoint = gk_spy
    orig_forward = net.forward

    def fwd_spy(x):
        y = orig_forward(x)
        rec.setdefault('net', []).append((x.detach().numpy().copy(), y.detach().numpy().copy()))
        return y
    net.forward = fwd_spy
    net64 = refmodel.SCNet(a)                       # the same module in float64: the "exact" output both float32 runs aim at
    net64.load_state_dict(sd)
    net64 = net64.double()

    P = synth.shipped_params('suncg')
    blob, names = {}, []
    for name, seed, tex_res, sig_feat in SCENES:
        rec.clear()
        data_s, data_t, R_gt = synth.make_room_scan_pair(seed, tex_res=tex_res)
        sf = P[:ALTER, 3] if sig_feat is None else np.full(ALTER, sig_feat)
        para = ru.opts(P[:ALTER, 0], P[:ALTER, 1], P[:ALTER, 2], sf)
        args = types.SimpleNamespace(snumclass=15, featureDim=32, outputType='rgbdnsf', maskMethod='second', alterStep=ALTER,
                                     dataset='suncg', para=para, representation='skybox', completion=True)
        np.random.seed(1000 + seed)
        R_hat = ref.RelativePoseEstimationViaCompletion(net, data_s, data_t, args)
        n_steps = len(rec.get('prim', []))
        print(name, "steps solved:", n_steps, "sift calls:", len(rec.get('sift', [])))
        names.append(name)
        blob[name + '/meta'] = np.array([seed, ALTER, n_steps, tex_res])
        blob[name + '/sigma'] = np.stack((P[:ALTER, 0], P[:ALTER, 1], P[:ALTER, 2], sf), 1)
        blob[name + '/R_gt'] = R_gt
        blob[name + '/R_final'] = R_hat
        first_run = {k: list(v) for k, v in rec.items()}
        for k in range(n_steps):
            dS, dT, para_k, T = rec['prim'][k]
            pre = "%s/step%d/" % (name, k)
            blob[pre + 'R_hat'] = T
            blob[pre + 'sift_s'], blob[pre + 'sift_t'] = rec['sift'][2 * k], rec['sift'][2 * k + 1]
            pts, ptsN, ptsW, ptt, pttN, pttW = rec['kp'][k]
            blob[pre + 'pts'], blob[pre + 'ptt'] = pts, ptt
            st = rec['rng'][k]
            blob[pre + 'rng_key'], blob[pre + 'rng_pos'] = st[1], np.array([st[2], st[3]])
            assert st[0] == 'MT19937' and st[3] == 0
            for side, d in (('s', dS), ('t', dT)):
                blob[pre + 'pc_' + side] = np.asarray(d['pc'])
                blob[pre + 'normal_' + side] = np.asarray(d['normal'])
                blob[pre + 'feat_' + side] = np.ascontiguousarray(d['feat'])
                blob[pre + 'weight_' + side] = np.asarray(d['weight'])
            blob[pre + 'feat_f_contig'] = np.array([int(np.asarray(dS['feat']).flags['C_CONTIGUOUS']), int(np.asarray(dT['feat']).flags['C_CONTIGUOUS'])])
            # oracle on the very same primitives: pins the restatement once more and yields the trace
            op = rp_oracle.Params(float(para_k.sigmaAngle1), float(para_k.sigmaAngle2), float(para_k.sigmaDist), float(para_k.sigmaFeat))
            tr = {}
            To = rp_oracle.solve_pair(dS, dT, op, tr)
            err = np.linalg.norm(To - T)
            print("  step %d: n_s=%d n_t=%d status=%s pairs=%s/%s  |oracle-ref|=%.2e  ang err vs gt %.2f deg  |t err| %.3f" % (
                k, len(dS['weight']), len(dT['weight']), tr.get('status'), tr.get('n_dist'), tr.get('n_angle'), err,
                float(ru.angular_distance_np(T[:3, :3], R_gt[:3, :3])[0]), np.linalg.norm(T[:3, 3] - R_gt[:3, 3])))
            assert err <= 1e-9
            blob[pre + 'topk'] = np.asarray(tr['topk'])
            blob[pre + 'counts'] = np.array([tr.get('n_dist', -1), tr.get('n_angle', -1), tr.get('status', -1)])
            x, y = rec['net'][k]
            blob[pre + 'net_sub'] = y[:, :, ::8, ::16].astype(np.float32)
            blob[pre + 'net_mean'] = y.mean(axis=(2, 3))
            blob[pre + 'net_std'] = y.std(axis=(2, 3))
            blob[pre + 'net_in_sum'] = x.astype(np.float64).sum(axis=(2, 3))
            with torch.no_grad():
                y64 = net64(torch.from_numpy(x).double()).numpy()
            blob[pre + 'net_sub32m64_x1e4'] = ((y.astype(np.float64) - y64)[:, :, ::8, ::16] * 1e4).astype(np.float16)   # float32 run minus float64 run
            print("    reference float32 vs float64 module: max-abs per head", [float(np.abs(y[:, a_:b_] - y64[:, a_:b_]).max()) for a_, b_ in ((0, 3), (3, 6), (6, 7), (7, 22), (22, 54))])
        # The alternation is chaotic with these (random) weights: warping scatters to rounded pixels and the bottleneck
        # BatchNorm sees two samples, so a pose that differs in the 5th digit gives a visibly different completion.  Measure
        # it on the reference itself: rerun with the pose of step 0 moved by 1e-5 rad / 1e-5 m and record how far the
        # reference's own later poses move.  (This is the yardstick for the free-running comparison in the GPU test.)
        rec.clear()
        calls = {'n': 0}

        def perturbed_rpe(*a_, **k_):
            T = orig_rpe(*a_, **k_)
            if calls['n'] == 0:
                c, s_ = np.cos(1e-5), np.sin(1e-5)
                dR = np.eye(4)
                dR[:3, :3] = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])
                dR[:3, 3] = 1e-5
                T = dR @ T
            calls['n'] += 1
            return T
        orig_rpe = ref.RelativePoseEstimation
        ref.RelativePoseEstimation = perturbed_rpe
        try:
            np.random.seed(1000 + seed)
            ref.RelativePoseEstimationViaCompletion(net, data_s, data_t, args)
        finally:
            ref.RelativePoseEstimation = orig_rpe
        sens = [float(np.linalg.norm(rec['prim'][k][3] - first_run['prim'][k][3])) for k in range(min(n_steps, len(rec.get('prim', []))))]
        print("  reference self-sensitivity (pose after step 0 moved by 1e-5): per-step |dT|_F =", sens)
        blob[name + '/ref_sensitivity_dT'] = np.array(sens)
    blob['names'] = np.array(names)
    path = os.path.join(HERE, 'via_completion_golden.npz')
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
