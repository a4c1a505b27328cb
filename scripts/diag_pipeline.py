import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import pipeline, synth, solver as S
from relativepose_b200.model.mymodel import Resnet18_8s
from relativepose_b200.RPModule.rputil import opts
B, K = 32, 103
rs = np.random.RandomState(0)
pts = np.stack((rs.uniform(1, 637, (2 * B, K)), rs.uniform(1, 157, (2 * B, K))), 2)
w = np.where((pts[..., 0] >= 160) & (pts[..., 0] <= 320), 1.0, 0.99)
nrm = rs.randn(2 * B, 160, 640, 3); nrm /= np.linalg.norm(nrm, axis=3, keepdims=True)
yy, xx = np.mgrid[0:160, 0:640]
depth = np.stack([2.5 + 1.5 * np.sin(xx / 37.0 + i) * np.cos(yy / 23.0) for i in range(2 * B)])
feat = torch.from_numpy(np.stack([synth.make_feature_map(i) for i in range(2 * B)])).cuda()
dep_d, nrm_d = torch.from_numpy(depth).cuda(), torch.from_numpy(nrm).cuda()
for sf in (0.05, 0.01):
    para = opts(*synth.shipped_params('matterport')[0]); para.sigmaFeat = sf
    d = pipeline.gather_primitives(feat, dep_d, nrm_d, pts, w, 'matterport')
    sol = S.default_solver('cuda:0')
    for _ in range(2): T, st, stats = sol.solve_device(d, [S.params_from_opts(para)])
    torch.cuda.synchronize(); t = time.perf_counter()
    T, st, stats = sol.solve_device(d, [S.params_from_opts(para)]); torch.cuda.synchronize()
    print("sigmaFeat %.2f: solve %.2f ms; status %s; stats mean %s" % (sf, (time.perf_counter() - t) * 1e3, np.bincount(st.cpu().numpy().clip(0)), stats.cpu().numpy().mean(0).round(1)))
