"""GPU parity for SCNet.forward: CUDA layers vs the oracle layer by layer, and the final
output vs the golden produced by the reference nn.Module itself (tests/golden/scnet_golden.npz).

Tolerances (float32 CUDA-core path; activations are O(1..10)):
  * output and every activation outside the bottleneck: max-abs 5e-4;
  * bottleneck (conv9 .. deconv6): max-abs 5e-3 -- conv9 is 1x1 spatial, so its BatchNorm batch is the 2 images of
    the pair: (x1-x2)/2/sqrt(((x1-x2)/2)^2+eps) amplifies float32 summation-order noise of a K=4608 dot product by up
    to 1/sqrt(eps)=316; the reference's own CPU and cuDNN paths disagree at this level.  The error decays again
    downstream (measured 1.4e-3 at conv9 -> 1.0e-4 at the output)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 5e-4
TOL_BOTTLENECK = 5e-3
BOTTLENECK = ('conv9', 'deconv9', 'deconv8', 'deconv7', 'deconv6')


def _args(snum, tanh):
    a = types.SimpleNamespace()
    a.batchnorm, a.useTanh, a.skipLayer, a.outputType, a.snumclass = 1, tanh, 1, 'rgbdnsf', snum
    return a


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scnet_golden.npz"))


@pytest.mark.parametrize("name", ["suncg", "scannet"])
def test_scnet_tensor_core_path(name):
    """tcgen05 path (bf16 operands, fp32 accumulate): same forward, documented looser tolerance.  bf16 has an 8-bit
    mantissa (2^-9 relative rounding per operand); through ~20 conv blocks with re-normalising BatchNorm the output
    differs from the float32 reference by ~1e-2 of its range.  Asserted: max-abs <= 0.25 and RMS <= 0.03 on outputs
    of magnitude ~10 (the bottleneck layers are excluded from any max-abs statement, see the fp32 test)."""
    import torch
    from oracle import scnet_oracle
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    from relativepose_b200.scnet_engine import ScnetEngine
    G = _golden()
    snum, tanh, seed, chk = G[name + '/meta']
    snum, tanh, seed = int(snum), int(tanh), int(seed)
    torch.manual_seed(0)
    net = SCNet(_args(snum, tanh))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.from_numpy(synth.make_panorama_pair(seed, str(G[name + '/dataset'])))
    net = net.cuda()
    tr = {}
    y = ScnetEngine(net, mode='tc').forward(x.cuda(), trace=tr)
    torch.cuda.synchronize()
    otr = {}
    with torch.no_grad():
        yo = scnet_oracle.forward_pair(sd, x, snum, bool(tanh), trace=otr)
    for k in otr:
        if k.endswith(':act') and k in tr:
            d = (tr[k].cpu() - otr[k])
            print("%-22s max %.3e rms %.3e (|ref|max %.3f)" % (k, d.abs().max().item(), d.pow(2).mean().sqrt().item(), otr[k].abs().max().item()))
    d = (y.cpu() - yo)
    emax, erms = d.abs().max().item(), d.pow(2).mean().sqrt().item()
    gsub = float(np.abs(y.cpu().numpy()[:, :, ::4, ::8] - G[name + '/sub']).max())
    print("tc final: max %.3e rms %.3e, vs reference golden (subsampled) max %.3e, |y|max %.2f" % (emax, erms, gsub, yo.abs().max().item()))
    assert emax <= 0.25 and erms <= 0.03 and gsub <= 0.25


@pytest.mark.parametrize("name", ["suncg", "scannet"])
def test_scnet_forward_matches_reference_golden(name):
    import torch
    from oracle import scnet_oracle
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    from relativepose_b200.scnet_engine import ScnetEngine
    G = _golden()
    snum, tanh, seed, chk = G[name + '/meta']
    snum, tanh, seed = int(snum), int(tanh), int(seed)
    torch.manual_seed(0)
    net = SCNet(_args(snum, tanh))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    got_chk = float(sum(v.double().abs().sum().item() for v in sd.values()))
    assert abs(got_chk - chk) <= 1e-6 * chk, "seeded weights differ from the ones the golden was made with"
    x = torch.from_numpy(synth.make_panorama_pair(seed, str(G[name + '/dataset'])))
    net = net.cuda()
    eng = ScnetEngine(net, mode='fp32')
    tr = {}
    y = eng.forward(x.cuda(), trace=tr)
    torch.cuda.synchronize()
    # layer-by-layer against the oracle (CPU fp32)
    otr = {}
    with torch.no_grad():
        yo = scnet_oracle.forward_pair(sd, x, snum, bool(tanh), trace=otr)
    worst = 0.0
    rows = []
    for k in otr:
        if k.endswith(':act') and k in tr:
            e = (tr[k].cpu() - otr[k]).abs().max().item()
            rows.append((k, e, otr[k].abs().max().item()))
            if k.split(':')[0] in BOTTLENECK:
                assert e <= TOL_BOTTLENECK, (k, e)
            else:
                worst = max(worst, e)
    e224 = (tr['out224'].cpu() - otr['out224']).abs().max().item()
    rows.append(('out224', e224, otr['out224'].abs().max().item()))
    for r in rows:
        print("%-22s err %.3e  (|ref|max %.3f)" % r)
    yc = y.cpu().numpy()
    err_or = float(np.abs(yc - yo.numpy()).max())
    err_gold = float(np.abs(yc[:, :, ::4, ::8] - G[name + '/sub']).max())
    print("final: vs oracle %.3e, vs reference golden %.3e" % (err_or, err_gold))
    assert worst <= TOL and e224 <= TOL
    assert err_or <= TOL and err_gold <= TOL
    assert np.abs(yc.mean(axis=(2, 3)) - G[name + '/mean']).max() <= TOL


def test_scnet_pairs_are_independent_bn_groups():
    """A batch of P pairs equals P separate forwards (the reference always forwards one pair: its BN batch)."""
    import torch
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    torch.manual_seed(0)
    net = SCNet(_args(15, 1)).cuda()          # default mode (tensor cores)
    xs = [torch.from_numpy(synth.make_panorama_pair(s, "suncg", 64, 256)).cuda() for s in (3, 4, 5)]
    yb = net(torch.cat(xs, 0))
    for i, x in enumerate(xs):
        yi = net(x)
        assert torch.equal(yb[2 * i:2 * i + 2], yi)


def test_scnet_module_surface():
    import torch
    from model.mymodel import SCNet
    torch.manual_seed(0)
    net = SCNet(_args(15, 1))
    keys = list(net.state_dict().keys())
    assert len(keys) == 103 and keys[0] == 'conv1rgb.0.weight' and 'deconv1f.bias' in keys
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 16, 32, 128))          # CPU tensor: no fallback


def _variants():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scnet_variants_golden.npz"))


@pytest.mark.parametrize("name", ["nobn", "noskip_sf", "partial_df", "nobn_noskip_f"])
@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_scnet_constructor_variants(name, mode):
    """batchnorm=0 (mymodel.py:23-27,35-39), skipLayer=0 (:333-357) and partial outputType against goldens from the
    unmodified reference class (tests/golden/make_scnet_variants_golden.py).  float32 path: 5e-4 of the output range
    (a batchnorm=0 network is not re-normalised, its outputs grow to O(100), hence relative); tensor-core path: the
    documented 16-bit tolerance (0.25 max-abs / 0.03 rms per unit of output range)."""
    import torch
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    from relativepose_b200.scnet_engine import ScnetEngine
    G = _variants()
    bn, skip, snum, tanh, seed, chk = G[name + '/meta']
    a = types.SimpleNamespace(batchnorm=int(bn), useTanh=int(tanh), skipLayer=int(skip), outputType=str(G[name + '/otype']),
                              snumclass=int(snum))
    torch.manual_seed(0)
    net = SCNet(a)
    got_chk = float(sum(v.double().abs().sum().item() for v in net.state_dict().values()))
    assert abs(got_chk - chk) <= 1e-6 * chk
    x = torch.from_numpy(synth.make_panorama_pair(int(seed), str(G[name + '/dataset'])))
    net = net.cuda()
    eng = ScnetEngine(net, mode=mode)
    ys = [eng.forward(x.cuda()) for _ in range(4)]            # eager, eager (recorded), plan replay, graph replay
    torch.cuda.synchronize()
    for y in ys[1:]:
        assert torch.equal(y, ys[0])
    yc = ys[0].cpu().numpy()
    assert yc.shape[1] == G[name + '/mean'].shape[1]
    sub = G[name + '/sub']
    rng_ = max(1.0, float(np.abs(sub).max()))
    d = yc[:, :, ::8, ::16] - sub
    emax, erms = float(np.abs(d).max()) / rng_, float(np.sqrt((d ** 2).mean())) / rng_
    emean = float(np.abs(yc.mean(axis=(2, 3)) - G[name + '/mean']).max()) / rng_
    print("%s %s: max %.3e rms %.3e mean %.3e (relative to |y|max %.2f)" % (name, mode, emax, erms, emean, rng_))
    if mode == 'fp32':
        assert emax <= TOL and emean <= TOL
    else:
        assert emax <= 0.25 and erms <= 0.03


@pytest.mark.parametrize("name", ["suncg", "scannet"])
def test_scnet_split_precision_tensor_core_mode(name):
    """mode='tc3': the tcgen05 kernels in split precision (half(x) w_hi + lo(x) w_hi + half(x) lo(w): one fused launch where the
    doubled halo fits shared memory, three launches for the stride-2 layers; float32 storage) against the reference module's golden.  Stated tolerance: 1e-3 max-abs on outputs of magnitude ~10
    (the float32 CUDA-core path is asserted at 5e-4, the 16-bit mode at 0.25); plan / graph replay bit-identical to eager."""
    import torch
    from oracle import scnet_oracle
    from relativepose_b200 import synth
    from relativepose_b200.model.mymodel import SCNet
    from relativepose_b200.scnet_engine import ScnetEngine
    G = _golden()
    snum, tanh, seed, chk = G[name + '/meta']
    snum, tanh, seed = int(snum), int(tanh), int(seed)
    torch.manual_seed(0)
    net = SCNet(_args(snum, tanh))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.from_numpy(synth.make_panorama_pair(seed, str(G[name + '/dataset'])))
    net = net.cuda()
    eng = ScnetEngine(net, mode='tc3')
    tr = {}
    y = eng.forward(x.cuda(), trace=tr)
    ys = [eng.forward(x.cuda()) for _ in range(4)]
    torch.cuda.synchronize()
    for y2 in ys:
        assert torch.equal(y2, y)
    otr = {}
    with torch.no_grad():
        yo = scnet_oracle.forward_pair(sd, x, snum, bool(tanh), trace=otr)
    for k in otr:
        if k.endswith(':act') and k in tr:
            d = (tr[k].cpu() - otr[k])
            print("%-22s max %.3e rms %.3e (|ref|max %.3f)" % (k, d.abs().max().item(), d.pow(2).mean().sqrt().item(), otr[k].abs().max().item()))
    d = (y.cpu() - yo)
    emax, erms = d.abs().max().item(), d.pow(2).mean().sqrt().item()
    ef = d[:, -32:].abs().max().item()
    gsub = float(np.abs(y.cpu().numpy()[:, :, ::4, ::8] - G[name + '/sub']).max())
    print("tc3 final: max %.3e rms %.3e (f head max %.3e), vs reference golden (subsampled) max %.3e, |y|max %.2f" % (emax, erms, ef, gsub, yo.abs().max().item()))
    assert emax <= 1e-3 and gsub <= 1e-3 and ef <= 4e-4


def test_scnet_head_subset_forward():
    """ScnetEngine.forward(heads=...) (what the batched alternation asks for: normals, depth, descriptors) computes the wanted
    heads bit-identically to the full forward and leaves the other decoder branches out (fewer launches)."""
    import torch
    from relativepose_b200 import _lib, synth
    from relativepose_b200.model.mymodel import SCNet
    from relativepose_b200.scnet_engine import ScnetEngine
    torch.manual_seed(0)
    net = SCNet(_args(15, 1)).cuda()
    x = torch.from_numpy(synth.make_panorama_pair(3, "suncg", 64, 256)).cuda()
    lib = _lib.load()
    full_eng, sub_eng = ScnetEngine(net), ScnetEngine(net)
    n0 = lib.rp_conv_launch_count()
    full = full_eng.forward(x)
    n1 = lib.rp_conv_launch_count()
    subs = [sub_eng.forward(x, heads=('n', 'd', 'f')) for _ in range(4)]          # eager, recorded, plan replay, graph replay
    torch.cuda.synchronize()
    n2 = lib.rp_conv_launch_count()
    want = list(range(3, 7)) + list(range(7 + 15, 7 + 15 + 32))                   # n (3), d (1), f (32) of 'rgbdnsf'
    for sub in subs:
        assert torch.equal(sub[:, want], full[:, want])
    assert (n2 - n1) < 4 * (n1 - n0)                                              # the rgb / s branches were not launched
