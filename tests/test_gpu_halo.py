"""GPU parity of the halo-tile tcgen05 convolution (csrc/scnet_halo.cu) layer by layer against torch's conv2d /
conv_transpose2d on the SAME bf16-rounded operands (producer BatchNorm + LeakyReLU applied in float32, rounded to
bf16; weights rounded to bf16) -- so the only difference left is the fp32 accumulation order (tolerance 2e-3 of the
output range; +2^-8 relative when the raw output is stored as bf16) -- and of the BN scale/shift the layer's
partial statistics produce."""
import numpy as np
import pytest

from relativepose_b200.scnet_engine import h16

pytestmark = pytest.mark.gpu

CASES = [
    # name, transposed, k, s, p, Hin, Win, [Cin...], Cout
    ("conv3x3s1", 0, 3, 1, 1, 40, 56, [64], 64),
    ("conv4x4s2_32to64", 0, 4, 2, 1, 112, 112, [32], 64),
    ("conv4x4s2_768to256", 0, 4, 2, 1, 56, 56, [768], 256),
    ("conv3x3s2p1", 0, 3, 2, 1, 40, 160, [64], 128),
    ("deconv4x4s2_cat", 1, 4, 2, 1, 28, 28, [128, 128], 64),
    ("deconv4x4s2_cat_c32", 1, 4, 2, 1, 56, 56, [64, 64], 32),
    ("deconv3x3s1", 1, 3, 1, 1, 17, 23, [64], 128),
    ("deconv3x3s2p0", 1, 3, 2, 0, 16, 15, [64], 32),
    ("conv3x3s1_tk16_stem", 0, 3, 1, 1, 48, 40, [16], 32),
    ("conv4x4s2_14to7", 0, 4, 2, 1, 14, 14, [512], 512),
    ("conv3x3p0_to1x1", 0, 3, 1, 0, 3, 3, [512], 1024),
    ("deconv3x3p0_from1x1", 1, 3, 1, 0, 1, 1, [1024], 512),
    ("deconv3x3s2p0_3to7", 1, 3, 2, 0, 3, 3, [512, 512], 512),
    ("conv3x3s1_many_tiles", 0, 3, 1, 1, 112, 112, [64], 64),
]

HEADS = [
    # name, Hin, Win, [Cin...], Cout, tanh
    ("head_rgb", 56, 40, [32, 32], 3, 0),
    ("head_sem", 56, 40, [64], 21, 0),
    ("head_feat_tanh", 56, 40, [64], 32, 1),
]


def _run(case, storage, flags=0):
    import torch
    import torch.nn.functional as F
    from relativepose_b200.scnet_engine import ScnetEngine, _Act
    name, tr, k, s, p, Hin, Win, cins, Cout = case
    dev = torch.device("cuda:0")
    G, gsz = 3, 2
    n = G * gsz
    Cin = sum(cins)
    g = torch.Generator(device="cpu").manual_seed(7)
    eng = ScnetEngine(None, mode='tc')
    eng.halo, eng.halo_flags = True, flags
    eng._P, eng._dev, eng._bufs = G, dev, {'partials': None}
    dt = h16() if storage == 'bf16' else torch.float32
    srcs, xs = [], []
    for c in cins:
        pitch = c + 8                                   # exercise pitch / channel offset
        raw = torch.randn((n, Hin, Win, pitch), generator=g).to(dev).to(dt)
        sc = (0.5 + torch.rand((G, pitch), generator=g)).to(dev)
        sh = (0.3 * torch.randn((G, pitch), generator=g)).to(dev)
        srcs.append(_Act(raw, Hin, Win, pitch, 8, c, sc, sh))
        xa = raw.float() * sc.repeat_interleave(gsz, 0)[:, None, None, :] + sh.repeat_interleave(gsz, 0)[:, None, None, :]
        xa = F.leaky_relu(xa, 0.1)[..., 8:8 + c].to(h16()).float()
        xs.append(xa)
    x = torch.cat(xs, 3).permute(0, 3, 1, 2).contiguous()
    if tr:
        w = torch.randn((Cin, Cout, k, k), generator=g).to(dev) / (Cin * k * k / (s * s)) ** 0.5
        Hout, Wout = (Hin - 1) * s - 2 * p + k, (Win - 1) * s - 2 * p + k
        wq = w.to(h16()).float()
        ref = F.conv_transpose2d(x.double(), wq.double(), stride=s, padding=p)
        eng._packed = {'L': w.permute(2, 3, 0, 1).contiguous()}
    else:
        w = torch.randn((Cout, Cin, k, k), generator=g).to(dev) / (Cin * k * k) ** 0.5
        Hout, Wout = (Hin + 2 * p - k) // s + 1, (Win + 2 * p - k) // s + 1
        wq = w.to(h16()).float()
        ref = F.conv2d(x.double(), wq.double(), stride=s, padding=p)
        eng._packed = {'L': w.permute(2, 3, 1, 0).contiguous()}
    opitch = Cout + 8
    obuf = torch.full((n, Hout, Wout, opitch), 768.0, device=dev).to(dt)
    out = _Act(obuf, Hout, Wout, opitch, 8, Cout, torch.zeros((G, opitch), device=dev), torch.zeros((G, opitch), device=dev))
    gamma = (0.5 + torch.rand(Cout, generator=g)).to(dev)
    beta = torch.randn(Cout, generator=g).to(dev)
    launches = []
    orig = eng.lib.rp_conv_layer_halo

    def spy(*a):                                         # the layer must really run on the halo kernel
        launches.append(1)
        return orig(*a)
    eng.lib.rp_conv_layer_halo = spy
    try:
        eng._conv('L', srcs, out, bool(tr), k, s, p, stream=torch.cuda.current_stream().cuda_stream, bn_params=(gamma, beta))
    finally:
        eng.lib.rp_conv_layer_halo = orig
    torch.cuda.synchronize()
    assert launches, "layer did not take the halo path"
    got = obuf[..., 8:8 + Cout].float().permute(0, 3, 1, 2).double()
    assert torch.all(obuf[..., :8].float() == 768.0), "wrote outside the channel window"
    rng = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    tol = 2e-3 * rng + (2.0 ** -8 * rng if storage == 'bf16' else 0.0)
    # BN scale/shift from the partial statistics vs the statistics of the stored tensor
    gg = got.reshape(G, gsz, Cout, -1).permute(0, 2, 1, 3).reshape(G, Cout, -1)
    mean, var = gg.mean(2), gg.var(2, unbiased=False)
    sc_ref = gamma.double() / torch.sqrt(var + 1e-5)
    sh_ref = beta.double() - mean * sc_ref
    e_sc = ((out.scale[:, 8:8 + Cout].double() - sc_ref).abs() / sc_ref.abs()).max().item()
    e_sh = (out.shift[:, 8:8 + Cout].double() - sh_ref).abs().max().item()
    return err, tol, rng, e_sc, e_sh


@pytest.mark.parametrize("storage", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_halo_layer_matches_torch(case, storage):
    err, tol, rng, e_sc, e_sh = _run(case, storage)
    print("%s/%s: max err %.3e (tol %.3e, range %.2f), bn scale rel %.2e shift abs %.2e" % (case[0], storage, err, tol, rng, e_sc, e_sh))
    assert err <= tol
    if case[5] * case[6] >= 64 or case[1]:      # (two-sample statistics of a 1x1 output amplify rounding by 1/sqrt(eps): checked end to end)
        assert e_sc <= 1e-3 and e_sh <= 5e-3


@pytest.mark.parametrize("case", [CASES[i] for i in (0, 3, 4, 11)], ids=[CASES[i][0] for i in (0, 3, 4, 11)])
def test_halo_layer_packed_half_batchnorm(case):
    """flags bit 1 (the engine's default for 16-bit sources): producer BatchNorm + LeakyReLU in packed half arithmetic.  The
    reference here applies them in float32 and rounds once, so scale / shift rounded to half show up: 2^-10 of the range more."""
    err, tol, rng, e_sc, e_sh = _run(case, "bf16", flags=2)
    print("%s/packed-half BN: max err %.3e (tol %.3e, range %.2f)" % (case[0], err, tol + 2.0 ** -10 * rng, rng))
    assert err <= tol + 2.0 ** -10 * rng and e_sc <= 2e-3


def test_halo_16bit_sources_arrive_by_tiled_tma():
    """16-bit sources can take the tiled-TMA loader (cp.async.bulk.tensor, one 5-D box per parity plane and K chunk; default
    for the single-tap layers, forced by flags bit 7); flags bit 6 selects the per-thread cp.async gather.  Both must give the
    same bits (same operands, same MMA order) -- stride-1, stride-2 (four parity planes) and transposed layers."""
    import torch
    from relativepose_b200 import _lib
    lib = _lib.load()
    n0 = lib.rp_conv_halo_tma_count()
    for case in (CASES[0], CASES[3], CASES[4]):
        a = _run(case, "bf16", flags=2 | 128)
        assert lib.rp_conv_halo_tma_count() > n0, "tensor-map encode failed: the layer fell back to the cp.async gather"
        n0 = lib.rp_conv_halo_tma_count()
        b = _run(case, "bf16", flags=2 | 64)
        assert lib.rp_conv_halo_tma_count() == n0
        assert a[0] == b[0], (case[0], a[0], b[0])       # identical max error against torch = identical output


def test_halo_layer_pitch16_variant():
    err, tol, rng, e_sc, e_sh = _run(CASES[0], "fp32", flags=1)
    assert err <= tol and e_sc <= 1e-3


@pytest.mark.parametrize("storage", ["fp32", "bf16"])
@pytest.mark.parametrize("case", HEADS, ids=[c[0] for c in HEADS])
def test_halo_1x1_head_with_bias(case, storage):
    """The 1x1 output heads (Conv2d + bias [+ tanh], no BatchNorm, float32 output at an unaligned channel offset of a
    wider tensor) on the halo kernel with Cout zero-padded to one n-tile."""
    import torch
    import torch.nn.functional as F
    from relativepose_b200.scnet_engine import ScnetEngine, _Act
    name, Hin, Win, cins, Cout, tanh = case
    dev = torch.device("cuda:0")
    G, gsz = 3, 2
    n = G * gsz
    g = torch.Generator(device="cpu").manual_seed(11)
    eng = ScnetEngine(None, mode='tc')
    eng.halo = True
    eng._P, eng._dev, eng._bufs = G, dev, {'partials': None}
    dt = h16() if storage == 'bf16' else torch.float32
    srcs, xs = [], []
    for c in cins:
        raw = torch.randn((n, Hin, Win, c), generator=g).to(dev).to(dt)
        sc = (0.5 + torch.rand((G, c), generator=g)).to(dev)
        sh = (0.3 * torch.randn((G, c), generator=g)).to(dev)
        srcs.append(_Act(raw, Hin, Win, c, 0, c, sc, sh))
        xa = raw.float() * sc.repeat_interleave(gsz, 0)[:, None, None, :] + sh.repeat_interleave(gsz, 0)[:, None, None, :]
        xa = F.leaky_relu(xa, 0.1)
        xs.append(xa.to(h16()).float() if Cout > 4 else xa)      # 3-channel heads run in float32 on CUDA cores
    x = torch.cat(xs, 3).permute(0, 3, 1, 2).contiguous()
    Cin = sum(cins)
    w = torch.randn((Cout, Cin, 1, 1), generator=g).to(dev) / Cin ** 0.5
    bias = torch.randn(Cout, generator=g).to(dev)
    ref = F.conv2d(x.double(), (w.to(h16()) if Cout > 4 else w).double(), bias.double())
    if tanh:
        ref = torch.tanh(ref)
    eng._packed = {'L': w.permute(2, 3, 1, 0).contiguous()}
    obuf = torch.full((n, Hin, Win, 54), 768.0, device=dev)
    out = _Act(obuf, Hin, Win, 54, 3, Cout)                       # channel offset 3: unaligned float4
    launches = []
    orig = eng.lib.rp_conv_layer_halo

    def spy(*a):
        launches.append(1)
        return orig(*a)
    eng.lib.rp_conv_layer_halo = spy
    try:
        eng._conv('L', srcs, out, False, 1, 1, 0, bn=False, bias=bias, tanh=bool(tanh), stream=torch.cuda.current_stream().cuda_stream)
    finally:
        eng.lib.rp_conv_layer_halo = orig
    torch.cuda.synchronize()
    assert bool(launches) == (Cout > 4), "wide heads take the halo kernel, 3-channel heads the CUDA-core head kernel"
    got = obuf[..., 3:3 + Cout].permute(0, 3, 1, 2).double()
    assert torch.all(obuf[..., :3] == 768.0) and torch.all(obuf[..., 3 + Cout:] == 768.0), "wrote outside the channel window"
    err = (got - ref).abs().max().item()
    print("%s/%s: max err %.3e (range %.2f)" % (name, storage, err, ref.abs().max().item()))
    assert err <= 4e-3 * max(1.0, ref.abs().max().item())   # a bf16 rounding tie of one activation (fmaf vs mul+add) moves an output by ~2e-3
