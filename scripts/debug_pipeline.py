import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import rp_oracle
from relativepose_b200 import synth
from relativepose_b200.model.mymodel import SCNet
from RPModule.rputil import opts, interpolate, getPixel
from relativepose_b200.solver import PoseSolver
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_solver import _debug_run
a = types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)
torch.manual_seed(0)
net = SCNet(a).cuda()
x = torch.from_numpy(synth.make_panorama_pair(21, "suncg")).cuda()
f = net(x)
i0 = 22
rs = np.random.RandomState(3)
def grid(n):
    p = np.stack((rs.uniform(1, 637, n), rs.uniform(1, 157, n)), 1)
    return p, p / np.array([640.0, 160.0]), np.where((p[:, 0] >= 160) & (p[:, 0] <= 320), 1.0, 0.99)
def data(k):
    depth = np.abs(f[k, 6].cpu().numpy().astype(np.float64)) + 0.5
    nrm = f[k, 3:6].cpu().numpy().astype(np.float64).transpose(1, 2, 0)
    nrm /= (np.linalg.norm(nrm, axis=2, keepdims=True) + 1e-12)
    return depth, nrm, f[k, i0:i0 + 32]
ps, psn, pw = grid(60); pt, ptn, tw = grid(55)
ds, ns_, fs = data(0); dt, nt_, ft = data(1)
p3s, nns = getPixel(ds, ns_, ps); p3t, nnt = getPixel(dt, nt_, pt)
dess = interpolate(fs, psn).cpu().numpy().T; dest = interpolate(ft, ptn).cpu().numpy().T
rec = {'pc_src': p3s.T, 'normal_src': nns, 'feat_src': dess, 'weight_src': pw, 'pc_tgt': p3t.T, 'normal_tgt': nnt, 'feat_tgt': dest, 'weight_tgt': tw}
for sf in (0.05, 0.2):
    P = synth.shipped_params('suncg')
    para = opts(*P[0]); para.sigmaFeat = sf
    op = rp_oracle.Params(*P[0]); op.sigmaFeat = sf
    out = _debug_run(PoseSolver("cuda:0"), [rec], para)
    tr = {}
    s = {'pc': rec['pc_src'], 'normal': rec['normal_src'], 'feat': rec['feat_src'], 'weight': rec['weight_src']}
    t = {'pc': rec['pc_tgt'], 'normal': rec['normal_tgt'], 'feat': rec['feat_tgt'], 'weight': rec['weight_tgt']}
    To = rp_oracle.solve_pair(s, t, op, tr)
    K = int(out['stats'][0, 7])
    got = np.sort(out['topk_idx'][:60, :K], axis=1)
    print("sigmaFeat", sf, "status gpu/oracle", out['status'][0], tr['status'], "stats", out['stats'][0])
    print("  topk rows equal:", int((got == tr['topk']).all(axis=1).sum()), "/ 60; zero rows:", int((tr['wij'].sum(1) == 0).sum()),
          "; dij equal:", np.array_equal(out['dij'][:60 * 55].reshape(60, 55), tr['dij']))
    print("  oracle n_dist, n_angle:", tr.get('n_dist'), tr.get('n_angle'), " |T-To| =", np.linalg.norm(out['T'][0] - To))
    if 'x' in tr:
        print("  oracle alternation count", len(tr['x']), "w range", tr['w'].min(), tr['w'].max())
