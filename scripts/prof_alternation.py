"""One RelativePoseEstimationViaCompletion_batch call (32 ScanNet-shape pairs, 3 alternation steps, host scans in) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum` (launch list of the whole step)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.argv = [sys.argv[0]]
import bench
from relativepose_b200 import pipeline, synth
from relativepose_b200.model.mymodel import SCNet
from relativepose_b200.RPModule.rputil import opts
B = 32
dev = torch.device("cuda:0")
rgb, nrm, depth, pts, w = bench.synth_scans(B)
torch.manual_seed(0)
net = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=1, outputType='rgbdnsf', snumclass=21)).to(dev)
P = synth.shipped_params('scannet')
pa = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
args = types.SimpleNamespace(snumclass=21, featureDim=32, outputType='rgbdnsf', maskMethod='kinect', alterStep=3,
                             dataset='scannet', para=pa, representation='skybox', completion=True)
for _ in range(3):
    pipeline.RelativePoseEstimationViaCompletion_batch(net, rgb, nrm, depth, pts, w, args)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipeline.RelativePoseEstimationViaCompletion_batch(net, rgb, nrm, depth, pts, w, args)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
