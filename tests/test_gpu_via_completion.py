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

# max-abs tolerance of the network output per head, [fp32 mode, tc (bf16 tensor-core) mode]; the f head is tanh bounded
# in (-1, 1), the others are unbounded regressions of magnitude ~10 with random weights
NET_TOL = {'fp32': {'rgb': 1e-3, 'n': 1e-3, 'd': 1e-3, 's': 1e-3, 'f': 2e-4},
           'tc': {'rgb': 0.30, 'n': 0.30, 'd': 0.30, 's': 0.30, 'f': 0.06}}
# pose tolerance ||T - T_ref||_F per teacher-forced step
POSE_TOL = {'fp32': 1e-4}


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
        steps.append({'net_sub': y[:, :, ::8, ::8].cpu().numpy()})
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
        assert r['kp_same_count'] and min(r['kp_rows_equal']) == 1.0, r      # identical keypoints
        assert r['feat_maxabs'] <= 2e-4 and r['pc_maxabs'] <= 2e-3, r
        assert r['topk_rows_equal'] >= 0.99, r
        assert r['dT'] <= POSE_TOL['fp32'], r


@pytest.mark.parametrize("name", ["room_a", "room_b"])
def test_via_completion_fp32_free_running(name):
    G = _golden()
    steps, T_final = _run(G, name, 'fp32', False)
    rows = _compare(G, name, 'fp32', steps, 'free')
    assert rows[0]['dT'] <= POSE_TOL['fp32']
    print("free-running final |T - T_ref|_F = %.3e" % np.linalg.norm(T_final - G[name + '/R_final']))
    assert np.linalg.norm(T_final - G[name + '/R_final']) <= 1e-3


@pytest.mark.parametrize("name", ["room_a", "room_b"])
def test_via_completion_tc_teacher_forced(name):
    """Default (bf16 tensor-core) network: per-head output error, fraction of keypoints / top-k rows that change, pose error.
    bf16 descriptors cannot meet 1e-4 on the pose (SURVEY.md 8d says so); what is asserted is the stated per-head bound
    and that the numbers are reported."""
    G = _golden()
    steps, _ = _run(G, name, 'tc', True)
    rows = _compare(G, name, 'tc', steps, 'teacher')
    for r in rows:
        for h, _, _ in HEADS:
            assert r['net_' + h] <= NET_TOL['tc'][h], (h, r)
        assert np.isfinite(r['dT'])
