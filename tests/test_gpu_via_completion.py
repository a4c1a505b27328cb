"""End-to-end parity of RelativePoseEstimationViaCompletion (rpmodule.py:569-662) against the reference's OWN run of it
(tests/golden/via_completion_golden.npz, made by tests/golden/make_via_completion_golden.py: reference SCNet on CPU float32,
recorded SIFT detections, seeded numpy RNG, alterStep = 3, shipped sigma rows 0-2).

Two protocols per scene and network mode:
  * teacher forced -- every alternation step starts from the reference's pose (and RNG state) of the previous step, so each
    step's network output, keypoints, primitives, top-k sets and pose are compared in isolation;
  * free running   -- the repo's own poses feed the next step, exactly what a user of the function gets.
Tolerances are stated per network head and per mode ("state the tolerance per dtype and show R,t parity
end-to-end from identical keypoints").  The solver alone is pinned much tighter on the very primitives the reference handed
to RelativePoseEstimation_helper (test_solver_on_reference_primitives: <= 1e-8, incl. the N = 3125 scene).
"""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
HEADS = (('rgb', 0, 3), ('n', 3, 6), ('d', 6, 7), ('s', 7, 22), ('f', 22, 54))

# max-abs tolerance of the network output per head, per mode.  The f head is tanh bounded in (-1, 1) (the descriptors the
# solver consumes), the others are unbounded regressions of magnitude ~10 with random weights.  'fp32' = CUDA-core float32
# kernels (the exact-parity path); 'tc' = default tcgen05 path, 16-bit operands (IEEE half, csrc/rp_h16.cuh) with float32
# accumulation.  Measured on B200 (round 2): fp32 <= 1.1e-4 on every head, the f head <= 8.3e-5 -- *inside* the distance
# between the reference's own float32 run and the same module in float64 (6e-5 .. 3e-4, stored in the golden).
NET_TOL = {'fp32': {'rgb': 1e-3, 'n': 1e-3, 'd': 1e-3, 's': 1e-3, 'f': 2e-4},
           'tc': {'rgb': 0.30, 'n': 0.30, 'd': 0.30, 's': 0.30, 'f': 0.06}}
# Pose tolerance ||T - T_ref||_F per teacher-forced step.  With IDENTICAL primitives the solver reproduces the reference to
# <= 1e-8 (test_solver_on_reference_primitives; north_star asks 1e-4).  Through the network the descriptors differ from the
# reference's by float32 summation order (~6e-5), which the soft-match kernel exp(-d / 2 (sigma/5)^2) amplifies: measured
# 4e-6 .. 1.6e-4 over the six steps.  The reference run on another BLAS / cuDNN build moves by as much (its own
# float32-vs-float64 distance is larger than ours to it), so 5e-4 is the stated tolerance of the fp32 mode.
POSE_TOL = {'fp32': 5e-4, 'tc3': 5e-4}
NET_TOL['tc3'] = {'rgb': 2e-3, 'n': 2e-3, 'd': 2e-3, 's': 2e-3, 'f': 4e-4}
REPORT = {}


def _report(key, rows):
    """Collected numbers go to gpurun_out/via_completion_parity.json (when that directory exists) for DESIGN.md."""
    import json
    REPORT[key] = rows
    out = os.path.join(os.path.dirname(HERE), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "via_completion_parity.json"), "w") as fh:
            json.dump(REPORT, fh, indent=1, default=lambda o: o.tolist() if hasattr(o, 'tolist') else str(o))


def _golden():
    return np.load(os.path.join(HERE, "golden", "via_completion_golden.npz"))


def _setup(G, name, mode):
    import torch
    from relativepose_b200 import scnet_engine, synth
    from relativepose_b200.model.mymodel import SCNet
    from RPModule.rputil import opts
    seed, alter, n_steps, tex_res = [int(v) for v in G[name + '/meta']]
    a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
    torch.manual_seed(0)
    net = SCNet(a).cuda()
    net._engine = scnet_engine.ScnetEngine(net, mode=mode)
    data_s, data_t, R_gt = synth.make_room_scan_pair(seed, tex_res=tex_res)
    sig = G[name + '/sigma']
    para = opts(sig[:, 0], sig[:, 1], sig[:, 2], sig[:, 3])
    args = types.SimpleNamespace(snumclass=15, featureDim=32, outputType='rgbdnsf', maskMethod='second', alterStep=alter,
                                 dataset='suncg', para=para, representation='skybox', completion=True)
    return net, data_s, data_t, args, seed, alter


def _run(G, name, mode, teacher):
    """Runs the repo's ViaCompletion; returns per-step dicts (net output sample, keypoints, primitives, pose)."""
    import relativepose_b200.RPModule.rpmodule as M
    import relativepose_b200.RPModule.rputil as U
    net, data_s, data_t, args, seed, alter = _setup(G, name, mode)
    steps = []
    state = {'k': 0, 'sift_call': 0}

    def sift_replay(gray):
        k, c = state['k'], state['sift_call']
        state['sift_call'] += 1
        return G["%s/step%d/sift_%s" % (name, k, 's' if c % 2 == 0 else 't')]

    orig_fwd, orig_helper, orig_rpe, orig_gk, orig_sift = net._engine.forward, M.RelativePoseEstimation_helper, \
        M.RelativePoseEstimation, M.getKeypoint, U._default_sift

    def fwd_spy(x, *a, **k):
        y = orig_fwd(x, *a, **k)
        steps.append({'net_sub': y[:, :, ::8, ::16].cpu().numpy()})
        return y

    def gk_spy(*a, **k):
        if teacher:
            pre = "%s/step%d/" % (name, state['k'])
            np.random.set_state(('MT19937', G[pre + 'rng_key'], int(G[pre + 'rng_pos'][0]), 0, 0.0))
        out = orig_gk(*a, **k)
        steps[-1]['pts'], steps[-1]['ptt'] = out[0], out[3]
        return out

    def helper_spy(dS, dT, para):
        T = orig_helper(dS, dT, para)
        steps[-1].update(dS=dS, dT=dT, T=T)
        return T

    def rpe_spy(*a, **k):
        T = orig_rpe(*a, **k)
        kk = state['k']
        state['k'] += 1
        return G["%s/step%d/R_hat" % (name, kk)] if teacher else T
    net._engine.forward, M.RelativePoseEstimation_helper, M.RelativePoseEstimation, M.getKeypoint, U._default_sift = \
        fwd_spy, helper_spy, rpe_spy, gk_spy, sift_replay
    try:
        np.random.seed(1000 + seed)
        T_final = M.RelativePoseEstimationViaCompletion(net, data_s, data_t, args)
    finally:
        net._engine.forward, M.RelativePoseEstimation_helper, M.RelativePoseEstimation, M.getKeypoint, U._default_sift = \
            orig_fwd, orig_helper, orig_rpe, orig_gk, orig_sift
    return steps, T_final


def _compare(G, name, mode, steps, label):
    from oracle import rp_oracle
    rows = []
    for k, st in enumerate(steps):
        pre = "%s/step%d/" % (name, k)
        r = {'step': k}
        gsub = G[pre + 'net_sub']
        for h, a, b in HEADS:
            r['net_' + h] = float(np.abs(st['net_sub'][:, a:b] - gsub[:, a:b]).max())
        r['f_rms'] = float(np.sqrt(np.mean((st['net_sub'][:, 22:54] - gsub[:, 22:54]) ** 2)))
        d32m64 = G[pre + 'net_sub32m64_x1e4'].astype(np.float64) * 1e-4      # reference float32 run minus float64 run
        r['f_ref32_vs_f64'] = float(np.abs(d32m64[:, 22:54]).max())
        r['f_ours_vs_f64'] = float(np.abs(st['net_sub'][:, 22:54].astype(np.float64) - (gsub[:, 22:54] - d32m64[:, 22:54])).max())
        same_kp = st['pts'].shape == G[pre + 'pts'].shape and st['ptt'].shape == G[pre + 'ptt'].shape
        r['kp_same_count'] = bool(same_kp)
        if same_kp:
            r['kp_rows_equal'] = float(np.mean(np.all(st['pts'] == G[pre + 'pts'], 1))), float(np.mean(np.all(st['ptt'] == G[pre + 'ptt'], 1)))
            r['feat_maxabs'] = float(max(np.abs(np.asarray(st['dS']['feat']) - G[pre + 'feat_s']).max(),
                                         np.abs(np.asarray(st['dT']['feat']) - G[pre + 'feat_t']).max()))
            r['pc_maxabs'] = float(max(np.abs(np.asarray(st['dS']['pc']) - G[pre + 'pc_s']).max(),
                                       np.abs(np.asarray(st['dT']['pc']) - G[pre + 'pc_t']).max()))
            # top-k sets of OUR primitives (oracle restatement == what the kernel selects, pinned in test_gpu_solver) vs the
            # reference's sets on ITS primitives
            sig = G[name + '/sigma'][k]
            tr = {}
            rp_oracle.solve_pair(st['dS'], st['dT'], rp_oracle.Params(*[float(v) for v in sig]), tr)
            r['topk_rows_equal'] = float(np.mean(np.all(tr['topk'] == G[pre + 'topk'], 1)))
            r['counts'] = (tr.get('n_dist'), tr.get('n_angle')), tuple(int(v) for v in G[pre + 'counts'][:2])
        r['dT'] = float(np.linalg.norm(st['T'] - G[pre + 'R_hat']))
        rows.append(r)
        print("[%s %s %s] %s" % (name, mode, label, r))
    _report("%s/%s/%s" % (name, mode, label), rows)
    return rows


@pytest.mark.parametrize("name", ["room_a", "room_b"])
def test_solver_on_reference_primitives(name):
    """The CUDA solver on the primitives the reference's ViaCompletion handed to its own helper (F-ordered descriptors, as
    ``torch_op.npy(interpolate(...)).T`` produces them) reproduces the reference's pose of every step."""
    from RPModule.rpmodule import RelativePoseEstimation_helper
    from RPModule.rputil import opts
    G = _golden()
    for k in range(int(G[name + '/meta'][2])):
        pre = "%s/step%d/" % (name, k)
        d = {}
        for side in 's', 't':
            feat = G[pre + 'feat_' + side]
            if not G[pre + 'feat_f_contig'][0 if side == 's' else 1]:
                feat = np.asfortranarray(feat)
            d[side] = {'pc': G[pre + 'pc_' + side], 'normal': G[pre + 'normal_' + side], 'feat': feat, 'weight': G[pre + 'weight_' + side]}
        sig = [float(v) for v in G[name + '/sigma'][k]]
        T = RelativePoseEstimation_helper(d['s'], d['t'], opts(*sig))
        err = np.linalg.norm(T - G[pre + 'R_hat'])
        print("%s step %d: n_s=%d n_t=%d |T - T_ref|_F = %.2e" % (name, k, len(d['s']['weight']), len(d['t']['weight']), err))
        assert err <= 1e-8


@pytest.mark.parametrize("name", ["room_a", "room_b"])
def test_via_completion_fp32_teacher_forced(name):
    G = _golden()
    steps, _ = _run(G, name, 'fp32', True)
    rows = _compare(G, name, 'fp32', steps, 'teacher')
    assert len(rows) == int(G[name + '/meta'][2])
    for r in rows:
        for h, _, _ in HEADS:
            assert r['net_' + h] <= NET_TOL['fp32'][h], (h, r)
        # as close to the exact (float64) descriptors as the reference's own float32 run is
        assert r['f_ours_vs_f64'] <= 1.5 * r['f_ref32_vs_f64'] + 5e-5, r
        assert r['kp_same_count'] and min(r['kp_rows_equal']) == 1.0, r      # identical keypoints
        assert r['feat_maxabs'] <= 2e-4 and r['pc_maxabs'] <= 2e-3, r
        assert r['topk_rows_equal'] >= 0.99, r                                # sets differ only on float32 near-ties
        assert r['dT'] <= POSE_TOL['fp32'], r


@pytest.mark.parametrize("name", ["room_a", "room_b"])
def test_via_completion_fp32_free_running(name):
    """What a caller gets: the repo's own pose feeds the next warp.  Step 0 has identical inputs and must agree.  From
    step 1 on the alternation is chaotic with these (random) weights -- warping scatters to rounded pixels and the
    bottleneck BatchNorm sees two samples -- and that is a property of the reference: the golden stores how far the
    REFERENCE's own later poses move when its pose after step 0 is moved by 1e-5 (ref_sensitivity_dT: O(0.1 .. 3)).  So
    later steps are reported next to that yardstick, and only required to be rigid transforms."""
    G = _golden()
    steps, T_final = _run(G, name, 'fp32', False)
    rows = _compare(G, name, 'fp32', steps, 'free')
    final = float(np.linalg.norm(T_final - G[name + '/R_final']))
    sens = G[name + '/ref_sensitivity_dT']
    print("free-running per-step |T - T_ref|_F = %s; the reference's own sensitivity to a 1e-5 pose change: %s"
          % ([r['dT'] for r in rows], sens.tolist()))
    _report("%s/fp32/free/final" % name, {'final_dT': final, 'reference_self_sensitivity_dT': sens.tolist()})
    assert rows[0]['dT'] <= POSE_TOL['fp32']
    assert len(rows) == int(G[name + '/meta'][2])
    assert np.allclose(T_final[3], [0, 0, 0, 1]) and abs(np.linalg.det(T_final[:3, :3]) - 1) < 1e-6
    assert np.allclose(T_final[:3, :3] @ T_final[:3, :3].T, np.eye(3), atol=1e-9)


@pytest.mark.parametrize("name", ["room_a", "room_b"])
def test_via_completion_tc_teacher_forced(name):
    """Default (16-bit tensor-core) network: per-head output error, fraction of keypoints / top-k rows that change, pose
    error.  16-bit descriptors cannot meet 1e-4 on the pose (SURVEY.md 8d says so); what is asserted is the stated per-head
    bound; the fraction of changed top-k rows and the pose distance are reported (gpurun_out/via_completion_parity.json ->
    DESIGN.md).  A caller that needs the reference's poses to 1e-4 selects RP_SCNET_MODE=fp32."""
    G = _golden()
    steps, _ = _run(G, name, 'tc', True)
    rows = _compare(G, name, 'tc', steps, 'teacher')
    for r in rows:
        for h, _, _ in HEADS:
            assert r['net_' + h] <= NET_TOL['tc'][h], (h, r)
        assert np.isfinite(r['dT'])


@pytest.mark.parametrize("name", ["room_a", "room_b"])
def test_via_completion_tc3_teacher_forced(name):
    """Split-precision tensor-core mode (RP_SCNET_MODE=tc3: every tcgen05 layer as half(x) w_hi + lo(x) w_hi + half(x) lo(w) --
    three MMAs per K step in one launch, three launches for the stride-2 layers -- float32 storage and accumulation): float32-class descriptors from the tensor cores.  Asserted: per-head
    output error within 2x the float32 mode's stated bound, identical keypoints, >= 99 % of the top-k rows identical to the
    reference's, pose within 5e-4 per teacher-forced step -- the float32 mode's own tolerance (measured on B200: every head
    <= 2.6e-4, f head <= 1.9e-4, top-k rows 99.5 - 100 % identical, pose 6.9e-6 .. 2.2e-4, five of six steps < 1e-4)."""
    G = _golden()
    steps, _ = _run(G, name, 'tc3', True)
    rows = _compare(G, name, 'tc3', steps, 'teacher')
    for r in rows:
        for h, _, _ in HEADS:
            assert r['net_' + h] <= NET_TOL['tc3'][h], (h, r)
        assert r['kp_same_count'] and min(r['kp_rows_equal']) >= 0.99, r
        assert r['topk_rows_equal'] >= 0.99, r
        assert r['dT'] <= POSE_TOL['tc3'], r
