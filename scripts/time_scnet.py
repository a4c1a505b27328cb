"""Time SCNet.forward on the GPU (CUDA events), P scan pairs per call."""
import sys, os, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relativepose_b200 import synth, _lib
from relativepose_b200.model.mymodel import SCNet
a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
torch.manual_seed(0)
net = SCNet(a).cuda()
print("config: RP_SCNET_HALO=%s RP_SCNET_ACT=%s RP_SCNET_MODE=%s" % (os.environ.get("RP_SCNET_HALO", "1"), os.environ.get("RP_SCNET_ACT", "bf16"), os.environ.get("RP_SCNET_MODE", "tc")))
for P in [int(v) for v in (sys.argv[1:] or ["1", "8"])]:
    x = torch.cat([torch.from_numpy(synth.make_panorama_pair(s, "suncg")) for s in range(P)], 0).cuda()
    for _ in range(5):          # includes the CUDA-graph capture on the third call
        y = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        y = net(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("P=%d pairs: %.3f ms per forward, %.3f ms per pair-step, %.1f TFLOP/s (72.27 GFLOP per pair-step)" % (
        P, ms, ms / P, 72.27e9 * P / (ms * 1e-3) / 1e12))
