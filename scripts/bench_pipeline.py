"""BASELINE configs[2] / [3]: network forward + RPModule for a batch of scan pairs, device resident (relativepose_b200/pipeline.py).

  configs[2]  B=32 Matterport-shape pairs, Resnet18_8s feature net (64 images per forward) + RPModule, n_s=n_t=103 (N=515)
  configs[3]  ScanNet-shape pairs, SCNet completion net + RPModule, 32 pairs per GPU (256 sharded over 8 GPUs), one
              alternation step and the full 3-step alternation (warp -> SCNet -> blend -> gather -> solve per step)

Keypoints are a seeded jittered grid (SIFT is OpenCV on the CPU, outside the path); random-init weights."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import pipeline, synth
from relativepose_b200.model.mymodel import SCNet, Resnet18_8s
from relativepose_b200.RPModule.rputil import opts

B, K = int(os.environ.get("PAIRS", 32)), 103
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
pts = np.stack((rs.uniform(1, 637, (2 * B, K)), rs.uniform(1, 157, (2 * B, K))), 2)
w = np.where((pts[..., 0] >= 160) & (pts[..., 0] <= 320), 1.0, 0.99)
nrm = rs.randn(2 * B, 160, 640, 3); nrm /= np.linalg.norm(nrm, axis=3, keepdims=True)
yy, xx = np.mgrid[0:160, 0:640]
depth = np.stack([2.5 + 1.5 * np.sin(xx / 37.0 + i) * np.cos(yy / 23.0) for i in range(2 * B)])
rgb = rs.uniform(0, 1, (2 * B, 160, 640, 3))


def timeit(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n


# ---- configs[2]: feature net + RPModule
torch.manual_seed(0)
fnet = Resnet18_8s(types.SimpleNamespace(num_input=7, useTanh=1)).cuda()
x7 = torch.cat((torch.from_numpy(rgb).float(), torch.from_numpy(nrm).float(), torch.from_numpy(depth).float().unsqueeze(3)), 3).permute(0, 3, 1, 2).contiguous().cuda()
dep_d, nrm_d = torch.from_numpy(depth).cuda(), torch.from_numpy(nrm).cuda()
para = opts(*synth.shipped_params('matterport')[0]); para.sigmaFeat = 0.05
pts_d, w_d = torch.from_numpy(pts).cuda(), torch.from_numpy(w).cuda()
def cfg2():
    f = fnet(x7)
    return pipeline.solve_from_maps(f, dep_d, nrm_d, pts_d, w_d, para, 'matterport')
t_net = timeit(lambda: fnet(x7)); t = timeit(cfg2)
f_ = fnet(x7)
t_solve = timeit(lambda: pipeline.solve_from_maps(f_, dep_d, nrm_d, pts_d, w_d, para, 'matterport'))
print("configs[2] B=%d Matterport pairs, Resnet18_8s (bf16 tcgen05) + RPModule N=%d: %.2f ms per batch = %.0f pairs/s (net %.2f ms, gather+solve+D2H %.2f ms)"
      % (B, 5 * K, t * 1e3, B / t, t_net * 1e3, t_solve * 1e3))

# ---- configs[3]: completion net + RPModule
torch.manual_seed(0)
cnet = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=1, outputType='rgbdnsf', snumclass=21)).cuda()
P = synth.shipped_params('scannet')
rgb_d = torch.from_numpy(rgb).float().cuda()
t_c = timeit(lambda: cnet(torch.zeros((2 * B, 16, 160, 640), device=dev)))
print("   (SCNet forward alone, %d pairs: %.2f ms)" % (B, t_c * 1e3))
for steps in (1, 3):
    pa = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
    args = types.SimpleNamespace(snumclass=21, featureDim=32, outputType='rgbdnsf', maskMethod='kinect', alterStep=steps,
                                 dataset='scannet', para=pa, representation='skybox', completion=True)
    t = timeit(lambda: pipeline.RelativePoseEstimationViaCompletion_batch(cnet, rgb_d, nrm_d, dep_d, pts_d, w_d, args), n=3, warm=3)
    print("configs[3] B=%d ScanNet pairs per GPU, SCNet (bf16 tcgen05) + RPModule N=%d, %d alternation step(s): %.2f ms per batch = %.0f pairs/s"
          % (B, 5 * K, steps, t * 1e3, B / t))
