"""One steady-state Resnet18_8s forward (64 images = configs[2]) between cudaProfilerStart/Stop."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relativepose_b200.model.mymodel import Resnet18_8s
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
net = Resnet18_8s(types.SimpleNamespace(num_input=7, useTanh=1)).cuda()
x = torch.randn(n, 7, 160, 640, device='cuda')
for _ in range(2):
    y = net(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = net(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
