"""Per-pair solver statistics of one alternation step (32 ScanNet-shape pairs): which pairs are slow and why."""
import os, sys, time, types, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.argv = [sys.argv[0]]
import bench
from relativepose_b200 import pipeline, synth, util as _util, solver as _solver
from relativepose_b200.model.mymodel import SCNet
from relativepose_b200.RPModule.rputil import opts
B = 32
dev = torch.device("cuda:0")
rgb, nrm, depth, pts, w = bench.synth_scans(B)
torch.manual_seed(0)
net = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=1, outputType='rgbdnsf', snumclass=21)).to(dev)
P = synth.shipped_params('scannet')
pa = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
rgb_d, nrm_d, dep_d = torch.from_numpy(rgb).to(dev), torch.from_numpy(nrm).to(dev), torch.from_numpy(depth).to(dev)
pts_d, w_d = torch.from_numpy(pts).to(dev), torch.from_numpy(w).to(dev)
n_img = 2 * B
full = torch.cat((rgb_d, nrm_d, dep_d.unsqueeze(3)), 3).permute(0, 3, 1, 2).contiguous()
vw, m, _g = _util.apply_mask(full, 'kinect')
views = torch.cat((vw, (vw[:, 6:7] != 0).float()), 1)
mask = m[:, 0].contiguous()
swap = (torch.arange(n_img, device=dev) ^ 1).to(torch.int32)
inp = torch.zeros((n_img, 16, 160, 640), dtype=torch.float32, device=dev)
inp[:, :8] = views
sv = _solver.default_solver(dev)
for label, R_hat in (("identity pose (step 0)", np.tile(np.eye(4), (B, 1, 1))), ("non-identity pose", np.tile(synth.make_pose(3), (B, 1, 1)))):
    Rs = np.empty((n_img, 4, 4)); Rs[0::2] = np.linalg.inv(R_hat); Rs[1::2] = R_hat
    _util.warping_device(inp, Rs, 'scannet', out=inp[:, 8:], src_index=swap)
    f = net(inp)
    nrm2, dep2 = _util.blend_completion_device(f, mask, nrm_d, dep_d)
    d = pipeline.gather_primitives(f[:, 28:60], dep2, nrm2, pts_d, w_d, 'scannet')
    para_this = copy.copy(pa)
    for name in ('sigmaAngle1', 'sigmaAngle2', 'sigmaDist', 'sigmaFeat'):
        setattr(para_this, name, getattr(pa, name)[0])
    pl = [_solver.params_from_opts(para_this)]
    for _ in range(2):
        T, st, stats = sv.solve_device(d, pl)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    T, st, stats = sv.solve_device(d, pl)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    s = stats.cpu().numpy(); stt = st.cpu().numpy()
    print(label, ": solve %.2f ms; status" % ms, np.bincount(stt - stt.min()), "min", stt.min())
    np.set_printoptions(linewidth=200)
    print("  columns N, M1, M2, NZ, tot_it, max_it, not_conv, K; all pairs:\n", s)
    print("  mean:", s.mean(0), " pairs solved by the accelerated iteration:", int(((s[:, 7] >> 8) & 1).sum()), "of", len(s))
    from relativepose_b200 import _lib
    for stage, nm in ((_lib.STAGE_TOPK, 'A only'), (_lib.STAGE_AFFINITY, 'A-D'), (_lib.STAGE_SOLVE, 'A-F')):
        for _ in range(2):
            sv.solve_device(d, pl, stop_after=stage)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sv.solve_device(d, pl, stop_after=stage)
        torch.cuda.synchronize(); print("   stop after %s: %.3f ms" % (nm, (time.perf_counter() - t0) * 1e3))
    print("  workspace key", sv._ws_key, "ws MB", sv._ws.numel() / 1e6)
