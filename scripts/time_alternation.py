"""Where one alternation step of RelativePoseEstimationViaCompletion_batch spends its time (CUDA events + host clock),
32 ScanNet-shape pairs, device-resident inputs."""
import os, sys, time, types
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
args = types.SimpleNamespace(snumclass=21, featureDim=32, outputType='rgbdnsf', maskMethod='kinect', alterStep=3,
                             dataset='scannet', para=pa, representation='skybox', completion=True)
rgb_d, nrm_d, dep_d = torch.from_numpy(rgb).to(dev), torch.from_numpy(nrm).to(dev), torch.from_numpy(depth).to(dev)
pts_d, w_d = torch.from_numpy(pts).to(dev), torch.from_numpy(w).to(dev)
for tag, a in (("host numpy scans in", (rgb, nrm, depth, pts, w)), ("device tensors in", (rgb_d, nrm_d, dep_d, pts_d, w_d))):
    for steps in (1, 3):
        args.alterStep = steps
        ms = bench.wall_ms(torch, lambda: pipeline.RelativePoseEstimationViaCompletion_batch(net, *a, args), 3, 3)
        print("%s, %d step(s): %.2f ms per 32 pairs" % (tag, steps, ms))
if hasattr(pipeline, "LAST_TIMING"):
    print("stage timing of the last call (ms):", pipeline.LAST_TIMING)

# ---- stage by stage (synchronising after each stage; one alternation step with a non-identity pose so the warp runs)
import copy
from relativepose_b200 import _lib
args.alterStep = 1
n_img = 2 * B
f32 = lambda a: torch.as_tensor(a, dtype=torch.float32).to(dev)
def stage_times():
    T = {}
    def tick(name, t0):
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    full = torch.cat((f32(rgb_d), f32(nrm_d), f32(dep_d).unsqueeze(3)), 3).permute(0, 3, 1, 2).contiguous()
    vw, m, _g = _util.apply_mask(full, args.maskMethod)
    views = torch.cat((vw, (vw[:, 6:7] != 0).float()), 1)
    mask = m[:, 0].contiguous()
    swap = (torch.arange(n_img, device=dev) ^ 1).to(torch.int32)
    inp = torch.empty((n_img, 16, 160, 640), dtype=torch.float32, device=dev)
    inp[:, :8] = views
    tick("setup (cat/mask/views)", t0)
    R_hat = np.tile(synth.make_pose(3), (B, 1, 1))
    t0 = time.perf_counter()
    Rs = np.empty((n_img, 4, 4)); Rs[0::2] = np.linalg.inv(R_hat); Rs[1::2] = R_hat
    _util.warping_device(inp, Rs, args.dataset, out=inp[:, 8:], src_index=swap)
    tick("warp", t0)
    t0 = time.perf_counter(); f = net(inp); tick("net", t0)
    t0 = time.perf_counter(); nrm2, dep2 = _util.blend_completion_device(f, mask, nrm_d, dep_d); tick("blend", t0)
    t0 = time.perf_counter(); d = pipeline.gather_primitives(f[:, 29:61], dep2, nrm2, pts_d, w_d, args.dataset); tick("gather", t0)
    para_this = copy.copy(pa)
    for name in ('sigmaAngle1', 'sigmaAngle2', 'sigmaDist', 'sigmaFeat'):
        setattr(para_this, name, getattr(pa, name)[0])
    sv = _solver.default_solver(dev)
    t0 = time.perf_counter(); Tt, st, _ = sv.solve_device_checked(d, [_solver.params_from_opts(para_this)]); tick("solve+status", t0)
    t0 = time.perf_counter(); Th = Tt.cpu().numpy(); tick("d2h", t0)
    return T
for _ in range(3):
    T = stage_times()
print("stage times (ms, each followed by a device synchronise):", {k: round(v, 2) for k, v in T.items()})
