"""GPU parity for Resnet18_8s.forward vs the golden from the reference class (stock-torchvision
shim, fork behaviour unpinned) and vs the oracle layer by layer.

Tolerances: float32 path max-abs 1e-3 on every block output and on the result (the BatchNorm batch here is the whole
call, no 2-sample bottleneck); tcgen05/bf16 path: max-abs <= 0.25, rms <= 0.03 on O(1..5) logits (see test_gpu_scnet)."""
import os
import types

import numpy as np
import pytest

from relativepose_b200.scnet_engine import h16

pytestmark = pytest.mark.gpu


def _golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resnet_golden.npz"))


@pytest.mark.parametrize("mode", ["fp32", "tc"])
@pytest.mark.parametrize("name", ["pair_tanh", "batch4_notanh"])
def test_resnet18_8s_forward(name, mode):
    import torch
    from oracle import resnet_oracle
    from relativepose_b200.model.mymodel import Resnet18_8s
    from relativepose_b200.resnet_engine import ResnetEngine
    G = _golden()
    tanh, n, seed, chk = G[name + '/meta']
    tanh, n, seed = int(tanh), int(n), int(seed)
    args = types.SimpleNamespace(num_input=7, useTanh=tanh)
    torch.manual_seed(0)
    net = Resnet18_8s(args)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    assert abs(float(sum(v.double().abs().sum().item() for v in sd.values())) - chk) <= 1e-6 * chk
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.uniform(-1, 1, size=(n, 7, 160, 640)).astype(np.float32))
    net = net.cuda()
    tr, otr = {}, {}
    y = ResnetEngine(net, mode=mode).forward(x.cuda(), trace=tr)
    torch.cuda.synchronize()
    yo = resnet_oracle.forward(sd, x, bool(tanh), trace=otr)
    worst = 0.0
    for k in otr:
        d = (tr[k].cpu() - otr[k])
        worst = max(worst, d.abs().max().item())
        print("%-28s max %.3e rms %.3e (|ref|max %.2f)" % (k, d.abs().max().item(), d.pow(2).mean().sqrt().item(), otr[k].abs().max().item()))
    d = y.cpu() - yo
    emax, erms = d.abs().max().item(), d.pow(2).mean().sqrt().item()
    gsub = float(np.abs(y.cpu().numpy()[:, :, ::4, ::8] - G[name + '/sub']).max())
    print("%s/%s final: max %.3e rms %.3e; vs reference golden %.3e" % (name, mode, emax, erms, gsub))
    if mode == "fp32":
        assert worst <= 1e-3 and emax <= 1e-3 and gsub <= 1e-3
    else:
        assert emax <= 0.25 and erms <= 0.03 and gsub <= 0.25


def test_resnet_module_surface():
    import torch
    from model.mymodel import Resnet18_8s
    net = Resnet18_8s(types.SimpleNamespace(num_input=7, useTanh=1))
    keys = list(net.state_dict().keys())
    assert 'resnet18_32s.layer2.0.downsample.0.weight' in keys and 'score_8s.bias' in keys and 'resnet18_32s.bn1.running_mean' in keys
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 7, 32, 128))


def test_im2col_bf16_matches_unfold():
    """rp_im2col_bf16 (the Resnet18_8s stem's patch matrix) against torch's unfold, K order (ky, kx, c), zero padded."""
    import torch
    import torch.nn.functional as F
    from relativepose_b200 import _lib
    lib = _lib.load()
    n, H, W, C, k, s, p = 3, 37, 70, 7, 7, 2, 3
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    Kp = -(-(k * k * C) // 32) * 32
    x = torch.randn((n, C, H, W), device='cuda')
    xin = x.permute(0, 2, 3, 1).contiguous()
    out = torch.full((n, Ho, Wo, Kp), 7.0, dtype=h16(), device='cuda')
    _lib.check(lib.rp_im2col_bf16(xin.data_ptr(), n, H, W, C, k, s, p, Ho, Wo, Kp, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "im2col")
    cols = F.unfold(x, k, padding=p, stride=s)                                   # [n, C*k*k, Ho*Wo], K order (c, ky, kx)
    ref = cols.view(n, C, k * k, Ho, Wo).permute(0, 3, 4, 2, 1).reshape(n, Ho, Wo, k * k * C)
    assert torch.equal(out[..., :k * k * C].float(), ref.to(h16()).float())
    assert torch.all(out[..., k * k * C:] == 0)
