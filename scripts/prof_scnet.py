"""One steady-state SCNet forward between cudaProfilerStart/Stop (use: ncu --profile-from-start off ...)."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relativepose_b200 import synth
from relativepose_b200.model.mymodel import SCNet
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
torch.manual_seed(0)
net = SCNet(a).cuda()
x = torch.cat([torch.from_numpy(synth.make_panorama_pair(s, "suncg")) for s in range(P)], 0).cuda()
for _ in range(2):
    y = net(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = net(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
