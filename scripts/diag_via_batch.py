"""Stage timing of one alternation step of RelativePoseEstimationViaCompletion_batch (32 ScanNet-shape pairs)."""
import os, sys, time, types, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import pipeline, synth, util as U, solver as S
from relativepose_b200.model.mymodel import SCNet
from relativepose_b200.RPModule.rputil import opts
B, K = 32, 103
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
pts = torch.from_numpy(np.stack((rs.uniform(1, 637, (2 * B, K)), rs.uniform(1, 157, (2 * B, K))), 2)).cuda()
w = torch.ones((2 * B, K), dtype=torch.float64, device=dev)
nrm = rs.randn(2 * B, 160, 640, 3); nrm /= np.linalg.norm(nrm, axis=3, keepdims=True)
yy, xx = np.mgrid[0:160, 0:640]
depth = np.stack([2.5 + 1.5 * np.sin(xx / 37.0 + i) * np.cos(yy / 23.0) for i in range(2 * B)])
nrm_d, dep_d = torch.from_numpy(nrm).cuda(), torch.from_numpy(depth).cuda()
views = torch.rand((2 * B, 8, 160, 640), device=dev)
mask = torch.zeros((2 * B, 160, 640), device=dev); mask[:, 47:113, 196:284] = 1
torch.manual_seed(0)
net = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=1, outputType='rgbdnsf', snumclass=21)).cuda()
P = synth.shipped_params('scannet')
para = opts(P[0, 0], P[0, 1], P[0, 2], 0.05)
swap = torch.arange(2 * B, device=dev) ^ 1
Rs = np.stack([synth.make_pose(i) for i in range(2 * B)])
sol = S.default_solver(dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(5):
    t0 = T(); warped = U.warping_device(views[swap], Rs, 'scannet')
    t1 = T(); x = torch.cat((views, warped), 1)
    t2 = T(); f = net(x)
    t3 = T(); n2, d2 = U.blend_completion_device(f, mask, nrm_d, dep_d)
    t4 = T(); d = pipeline.gather_primitives(f[:, 29:61], d2, n2, pts, w, 'scannet')
    t5 = T(); Tt, st, _ = sol.solve_device(d, [S.params_from_opts(para)])
    t6 = T(); R = Tt.cpu().numpy(); Ri = np.linalg.inv(R)
    t7 = T()
    if it >= 3:
        print("warp %.2f  cat %.2f  scnet %.2f  blend %.2f  gather %.2f  solve %.2f  d2h+inv %.2f  total %.2f ms" % tuple(
            1e3 * v for v in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t7 - t0)))
