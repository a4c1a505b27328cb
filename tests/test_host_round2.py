"""Host-side logic added in round 2 (no GPU needed): util.apply_mask against the reference's own, the reference install used
by bench.py's CPU arm, the rendered room scenes, the solver's status policy and workspace sizing rules."""
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_util():
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference tree not present")
    import torch
    ref = ref_loader.load_reference_rpmodule()
    keep = lambda var, cuda=True, volatile=False: (torch.from_numpy(var).float() if isinstance(var, np.ndarray) else var.float())  # noqa: E731
    ref.util.torch_op.v = keep
    return ref.util


@pytest.mark.parametrize("method", ["second", "kinect", "something-else"])
def test_apply_mask_matches_reference(method):
    """util.apply_mask (util.py:209-232): same 3-tuple (masked x, mask, geow) as the reference, for both shipped mask
    methods and for an unknown one (mask and geow stay zero)."""
    import torch
    from relativepose_b200.util import apply_mask
    util = _reference_util()
    x = torch.rand(2, 8, 160, 640)
    xr, mr, gr = util.apply_mask(x.clone(), method)
    xm, mm, gm = apply_mask(x.clone(), method, "extra-positional-arguments-are-accepted")
    assert torch.equal(xm, xr.float()) and torch.equal(mm, mr.float())
    gr = torch.as_tensor(np.asarray(gr), dtype=torch.float32) if not torch.is_tensor(gr) else gr.float()
    assert gm.shape == gr.shape and torch.allclose(gm, gr, atol=1e-7)


def test_reference_install_is_the_unmodified_tree():
    """oracle/install_reference.py: the files under baseline/_ref are byte-identical to /root/reference (when both exist)."""
    src, dst = "/root/reference", os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(src):
        pytest.skip("reference tree not present")
    sys.path.insert(0, ROOT)
    from oracle import install_reference
    assert install_reference.install() in ("present", "installed")
    for rel in ("RPModule/rpmodule.py", "RPModule/rputil.py", "util.py", "config.py", "utils/torch_op.py", "model/mymodel.py"):
        assert open(os.path.join(src, rel), "rb").read() == open(os.path.join(dst, rel), "rb").read(), rel


def test_room_scan_pair_is_geometrically_consistent():
    """synth.make_room_scan_pair: the two skybox scans see the same walls -- source points moved by R_gt land on the target's
    point cloud (oracle Pano2PointCloud, util.py:751-773) within the pixel spacing."""
    from scipy.spatial import cKDTree
    from oracle import warp_oracle
    from relativepose_b200 import synth
    s, t, R = synth.make_room_scan_pair(4, tex_res=12)
    assert s['rgb'].shape == (160, 640, 3) and s['depth'].min() > 0.5
    assert np.allclose(np.linalg.norm(s['norm'], axis=2), 1.0, atol=1e-12)
    assert np.allclose(R[:3, :3] @ R[:3, :3].T, np.eye(3), atol=1e-12)
    ps = warp_oracle.pano2pointcloud(s['depth'], 'suncg')
    pt = warp_oracle.pano2pointcloud(t['depth'], 'suncg')
    moved = (R[:3, :3] @ ps + R[:3, 3:4]).T[::97]
    d, _ = cKDTree(pt.T).query(moved)
    assert np.median(d) < 0.03           # pixel spacing at 3 m is ~4 cm


def test_status_policy():
    """PoseSolver.check_status: overflow of the bounded candidate list -> one redo at full capacity; unsupported -> raise."""
    from relativepose_b200 import _lib
    from relativepose_b200.solver import PoseSolver
    chk = PoseSolver.check_status
    calls = []

    def redo_ok():
        calls.append(1)
        return np.zeros(3, np.int32)
    assert chk(None, np.array([0, 1, 5], np.int32), redo_ok) is False and not calls
    assert chk(None, np.array([0, _lib.STATUS_EDGE_OVERFLOW, 0], np.int32), redo_ok) is True and len(calls) == 1
    with pytest.raises(RuntimeError):
        chk(None, np.array([_lib.STATUS_UNSUPPORTED], np.int32), redo_ok)
    with pytest.raises(RuntimeError):
        chk(None, np.array([_lib.STATUS_EDGE_OVERFLOW], np.int32), lambda: np.array([_lib.STATUS_EDGE_OVERFLOW], np.int32))


def test_bounded_edge_capacity_rule():
    from relativepose_b200.solver import PoseSolver
    s = types.SimpleNamespace(edge_frac=None, AUTO_EDGE_FRAC=PoseSolver.AUTO_EDGE_FRAC, AUTO_EDGE_FLOOR=PoseSolver.AUTO_EDGE_FLOOR)
    cap = lambda n_s: PoseSolver._edge_cap(s, n_s, 5)                                       # noqa: E731
    assert cap(20) == 0                                   # tiny pair: the worst case is below the floor -> full capacity
    P515 = 515 * 514 // 2
    assert cap(103) == 65536 < P515                       # N = 515: the 64 Ki floor
    N = 625 * 5
    assert cap(625) == int(N * (N - 1) // 2 * 0.25)       # N = 3125: a quarter of the worst case
    s.edge_frac = 1.0
    assert cap(103) == 0


def test_packed_batch_layouts_and_sum_order():
    """PackedBatch: concatenation is exact for C-ordered, F-ordered and list inputs; the per-pair float32 summation-order
    flag follows the memory layout of the descriptor arrays (rpmodule.py:531-532 hands over transposed views)."""
    from relativepose_b200 import synth
    from relativepose_b200.solver import PackedBatch
    recs = [synth.make_pair(i, 10 + i, 12 + i) for i in range(300)]          # >= 256: the threaded fill path
    r1 = dict(recs[1])
    r1['feat_src'] = np.asfortranarray(r1['feat_src'])
    r1['weight_tgt'] = list(r1['weight_tgt'])
    recs[1] = r1
    pk = PackedBatch(recs, pin=False)
    assert pk.B == 300 and pk.sum_order_t.tolist()[:3] == [0, 1, 0]
    for b in (0, 1, 150, 299):
        lo, hi = int(pk.off_s[b]), int(pk.off_s[b + 1])
        assert np.array_equal(pk.pc_s.numpy()[lo:hi], recs[b]['pc_src']) and np.array_equal(pk.feat_s.numpy()[lo:hi], recs[b]['feat_src'])
        lo, hi = int(pk.off_t[b]), int(pk.off_t[b + 1])
        assert np.array_equal(pk.w_t.numpy()[lo:hi], np.asarray(recs[b]['weight_tgt'])) and np.array_equal(pk.nrm_t.numpy()[lo:hi], recs[b]['normal_tgt'])


def test_scnet_constructor_variants_state_dict_and_oracle():
    """SCNet(batchnorm=0 / skipLayer=0 / partial outputType): parameter names follow the reference's layer table
    (mymodel.py:151-231) and the oracle restatement reproduces the reference's golden for one variant."""
    import os
    import types
    import numpy as np
    import torch
    from oracle import scnet_oracle
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scnet_variants_golden.npz"))
    a = types.SimpleNamespace(batchnorm=0, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
    keys = list(SCNet(a).state_dict().keys())
    assert 'conv4.0.bias' in keys and not any('.1.' in k for k in keys)
    a = types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=0, outputType='sf', snumclass=21)
    net = SCNet(a)
    assert net.deconv8[0].weight.shape[0] == 512 and not hasattr(net, 'deconv3rgb')
    assert [h for h, _ in net.head_channels()] == ['s', 'f'] and net.head_channels()[0][1] == 21
    import pytest
    with pytest.raises(NotImplementedError):
        SCNet(types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=0, outputType='rgbdnsf', snumclass=15))
    name = 'nobn_noskip_f'
    bn, skip, snum, tanh, seed, chk = G[name + '/meta']
    a = types.SimpleNamespace(batchnorm=int(bn), useTanh=int(tanh), skipLayer=int(skip), outputType=str(G[name + '/otype']),
                              snumclass=int(snum))
    torch.manual_seed(0)
    net = SCNet(a)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.from_numpy(synth.make_panorama_pair(int(seed), str(G[name + '/dataset'])))
    torch.set_num_threads(8)
    y = scnet_oracle.forward(sd, x, int(snum), bool(tanh), skip=bool(skip), heads=tuple(net.heads)).numpy()
    sub = G[name + '/sub']
    assert np.abs(y[:, :, ::8, ::16] - sub).max() <= 1e-4 * max(1.0, np.abs(sub).max())


def test_resnet_stem_space_to_depth_weights():
    """The Resnet18_8s stem (mymodel.py:51-54,85: conv 7x7, stride 2, padding 3) as a 4x4 stride-1 convolution over the 2x2
    space-to-depth input: the weight rearrangement of resnet_engine.stem_weights_s2d against torch's conv2d on the CPU."""
    import torch
    import torch.nn.functional as F
    from relativepose_b200.resnet_engine import stem_weights_s2d
    torch.manual_seed(3)
    n, C, H, W, Co = 2, 7, 20, 36, 5
    x = torch.randn(n, C, H, W, dtype=torch.float64)
    w = torch.randn(Co, C, 7, 7, dtype=torch.float64)
    ref = F.conv2d(x, w, None, stride=2, padding=3)                                   # [n,Co,10,18]
    w4 = stem_weights_s2d(w.permute(2, 3, 1, 0).float(), 32).double()                 # [4,4,32,Co]
    # space-to-depth exactly as the kernel lays it out: channel (dy*2+dx)*C + c
    sd = torch.zeros(n, 32, H // 2, W // 2, dtype=torch.float64)
    for dy in range(2):
        for dx in range(2):
            q = dy * 2 + dx
            sd[:, q * C:(q + 1) * C] = x[:, :, dy::2, dx::2]
    out = F.conv2d(F.pad(sd, (2, 1, 2, 1)), w4.permute(3, 2, 0, 1), None, stride=1)   # padding 2 before, 1 after = rows 0..H/2-1 of p = 2
    assert out.shape == ref.shape
    assert (out - ref).abs().max() <= 1e-5 * ref.abs().max()


def test_wave_aligned_chunks():
    """PoseSolver._solve_pipelined cuts a batch into waves of resident CTAs (solver.wave_chunks): whole waves, in order, covering
    every pair exactly once; a short tail joins the previous chunk."""
    from relativepose_b200.solver import wave_chunks
    assert wave_chunks(4096, 592) == [0, 592, 1184, 1776, 2368, 2960, 3552, 4096]
    assert wave_chunks(592, 592) == [0, 592]
    assert wave_chunks(100, 592) == [0, 100]
    assert wave_chunks(1200, 592) == [0, 592, 1200]                # 16-pair tail joins the second wave
    assert wave_chunks(1400, 592) == [0, 592, 1184, 1400]
    for B, per in ((1, 1), (7, 3), (5000, 148), (2048, 592)):
        b = wave_chunks(B, per)
        assert b[0] == 0 and b[-1] == B and all(lo < hi for lo, hi in zip(b, b[1:]))


def test_record_packing_extension_matches_numpy():
    """PackedBatch through the C packing extension (csrc/rp_pack.c: buffer protocol + memcpy with the GIL released) equals the
    numpy.concatenate path array for array -- incl. a Fortran-ordered descriptor array, a negative-stride view and a record that
    needs a dtype conversion (the extension declines that field, numpy converts)."""
    from relativepose_b200 import solver, synth
    from relativepose_b200.solver import PackedBatch
    ext = solver._pack_ext()
    assert ext is not None, "relativepose_b200/_rp_pack.so is missing: python -m relativepose_b200.build"
    recs = [dict(r) for r in synth.make_batch(77, 300, 40)]
    recs[5]['feat_src'] = np.asfortranarray(recs[5]['feat_src'])
    recs[7]['pc_tgt'] = recs[7]['pc_tgt'][::-1]
    recs[9]['normal_src'] = recs[9]['normal_src'].astype(np.float32)
    a = PackedBatch(recs, pin=False)
    try:
        solver._PACK_EXT[0] = None                                   # numpy path
        b = PackedBatch(recs, pin=False)
    finally:
        solver._PACK_EXT[0], solver._PACK_EXT[1] = None, False       # re-probe on next use
    for f in PackedBatch.FIELDS + ("off_s_t", "off_t_t", "sum_order_t"):
        assert np.array_equal(getattr(a, f).numpy(), getattr(b, f).numpy()), f
    assert int(a.sum_order_t.numpy()[5]) == 1 and int(a.sum_order_t.numpy()[6]) == 0
    # direct calls: row counts, declined cases
    dst = np.empty((300 * 40, 3), np.float64)
    cnt = np.frombuffer(ext.pack_field(recs, 'pc_src', dst, 'd', 8, 3), dtype=np.int64)
    assert cnt.tolist() == [40] * 300 and np.array_equal(dst, np.concatenate([r['pc_src'] for r in recs], 0))
    assert ext.pack_field(recs, 'normal_src', dst, 'd', 8, 3) is None            # record 9 is float32
    assert ext.pack_field(recs, 'no_such_key', dst, 'd', 8, 3) is None
    assert ext.pack_field(recs, 'pc_src', np.empty((10, 3)), 'd', 8, 3) is None   # destination too small
