"""Time Resnet18_8s.forward on the GPU (CUDA events), 64 images = configs[2]."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.argv = [sys.argv[0]]
import bench
from relativepose_b200.model.mymodel import Resnet18_8s
torch.manual_seed(0)
net = Resnet18_8s(types.SimpleNamespace(num_input=7, useTanh=1)).cuda()
x = torch.randn(64, 7, 160, 640, device='cuda')
print("Resnet18_8s 64 images: %.3f ms" % bench.device_time_ms(torch, lambda: net(x), 10, 5))
