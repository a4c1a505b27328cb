"""3-step alternation, 32 ScanNet-shape pairs, host scans in: time against the network chunk size."""
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
print("dtypes", rgb.dtype, nrm.dtype, depth.dtype)
torch.manual_seed(0)
net = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=1, outputType='rgbdnsf', snumclass=21)).to(dev)
P = synth.shipped_params('scannet')
pa = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
args = types.SimpleNamespace(snumclass=21, featureDim=32, outputType='rgbdnsf', maskMethod='kinect', alterStep=3,
                             dataset='scannet', para=pa, representation='skybox', completion=True)
ref = None
for chunk in (32, 16, 8, 32):
    fn = lambda: pipeline.RelativePoseEstimationViaCompletion_batch(net, rgb, nrm, depth, pts, w, args, chunk=chunk)
    ms, mx = bench.wall_ms_median(torch, fn, 5, 4)
    T = fn()
    if ref is None:
        ref = T
    print("chunk %2d: %.2f ms median (%.2f slowest); bitwise equal to chunk 32: %s" % (chunk, ms, mx, bool(np.array_equal(T, ref))))
rgb32 = rgb.astype(np.float32)
for chunk in (32, 16):
    fn = lambda: pipeline.RelativePoseEstimationViaCompletion_batch(net, rgb32, nrm, depth, pts, w, args, chunk=chunk)
    ms, mx = bench.wall_ms_median(torch, fn, 5, 4)
    print("float32 rgb, chunk %2d: %.2f ms median" % (chunk, ms))
