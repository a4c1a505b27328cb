"""The network forward as one native call (rp_scnet_forward / rp_resnet18_8s_forward): after a
warm-up run the layer calls are frozen into an op list; replaying it (directly or through the CUDA graph that captures it)
must reproduce the eager layer-by-layer forward bit for bit."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spy(lib, name, calls):
    orig = getattr(lib, name)

    def f(*a):
        calls.append(name)
        return orig(*a)
    setattr(lib, name, f)
    return orig


@pytest.mark.parametrize("graph", [False, True])
def test_scnet_forward_plan_equals_eager(graph):
    import torch
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    from relativepose_b200.scnet_engine import ScnetEngine
    torch.manual_seed(0)
    net = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)).cuda()
    xs = [torch.from_numpy(synth.make_panorama_pair(s, "suncg")).cuda() for s in (1, 2)]
    eng = ScnetEngine(net)
    eng.use_graph = graph
    ref = [eng._forward_eager(x).clone() for x in xs]
    calls = []
    orig = _spy(eng.lib, "rp_scnet_forward", calls)
    try:
        outs = [eng.forward(xs[i % 2]) for i in range(6)]          # eager, record, replay x4 (graph capture on the third)
    finally:
        eng.lib.rp_scnet_forward = orig
    for i, y in enumerate(outs):
        assert torch.equal(y, ref[i % 2]), "call %d differs from the eager forward" % i
    assert len(calls) >= (1 if graph else 4), calls


def test_resnet_forward_plan_equals_eager():
    import torch
    from relativepose_b200.model.mymodel import Resnet18_8s
    from relativepose_b200.resnet_engine import ResnetEngine
    torch.manual_seed(0)
    net = Resnet18_8s(types.SimpleNamespace(num_input=7, useTanh=1)).cuda()
    g = torch.Generator(device='cpu').manual_seed(3)
    xs = [torch.randn((4, 7, 160, 640), generator=g).cuda() for _ in range(2)]
    eng = ResnetEngine(net)
    ref = [eng._forward_impl(x).clone() for x in xs]
    calls = []
    orig = _spy(eng.lib, "rp_resnet18_8s_forward", calls)
    try:
        outs = [eng.forward(xs[i % 2]) for i in range(5)]
    finally:
        eng.lib.rp_resnet18_8s_forward = orig
    for i, y in enumerate(outs):
        assert torch.equal(y, ref[i % 2]), "call %d differs from the eager forward" % i
    assert len(calls) == 3


def test_affinity_build_stage_entry():
    """rp_affinity_build (rpmodule.py:342-472) returns the same surviving pairs and weights as the oracle's stage trace."""
    import ctypes
    import torch
    from oracle import rp_oracle
    from relativepose_b200 import _lib, synth
    from relativepose_b200.RPModule.rputil import opts
    from relativepose_b200.solver import PackedBatch, default_solver, params_from_opts
    rec = synth.make_pair(77, 40, 36)
    para = opts(*synth.shipped_params('suncg')[0])
    s, t = synth.record_to_dicts(rec)
    tr = {}
    rp_oracle.solve_pair(s, t, rp_oracle.Params(*synth.shipped_params('suncg')[0]), tr)
    sol = default_solver('cuda:0')
    d = PackedBatch([rec]).to_device(sol.device)
    lib = _lib.load()
    plist = [params_from_opts(para)]
    K = 5
    ws, key = sol._workspace(1, d.max_ns, d.max_nt, K, d.feat_dim, 0)
    par = sol._params_device(plist)
    cap = 4096
    dev = sol.device
    topk = torch.full((40, K), -1, dtype=torch.int32, device=dev)
    erc = torch.zeros((1, cap, 2), dtype=torch.int32, device=dev)
    ew = torch.zeros((1, cap), dtype=torch.float64, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    stats = torch.zeros((1, 8), dtype=torch.int32, device=dev)
    zr = d.zero_rows(K, K, dev)
    rc = lib.rp_affinity_build(1, d.off_s_t.data_ptr(), d.off_t_t.data_ptr(), d.pc_s.data_ptr(), d.nrm_s.data_ptr(), d.feat_s.data_ptr(),
                               d.w_s.data_ptr(), d.pc_t.data_ptr(), d.nrm_t.data_ptr(), d.feat_t.data_ptr(), d.w_t.data_ptr(), d.feat_dim,
                               par.data_ptr(), None, zr.data_ptr(), d.sum_order_t.data_ptr(), key[0], key[1], K, key[5], cap,
                               ws.data_ptr(), ws.numel(), topk.data_ptr(), erc.data_ptr(), ew.data_ptr(), status.data_ptr(),
                               stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    M = int(stats[0, 2])
    assert M == len(tr['w']) and int(status[0]) == 0
    got = np.sort(erc[0, :M].cpu().numpy(), axis=1)
    ref = np.sort(np.asarray(tr['pairs']), axis=1)
    og, orf = np.lexsort((got[:, 1], got[:, 0])), np.lexsort((ref[:, 1], ref[:, 0]))
    assert np.array_equal(got[og], ref[orf])
    assert np.allclose(ew[0, :M].cpu().numpy()[og], np.asarray(tr['w'])[orf], rtol=1e-11, atol=0)
