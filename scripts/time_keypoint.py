"""Time the keypoint-augmentation kernels (rputil.match_sample) and the view warp on the GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import synth, util
from relativepose_b200.RPModule import rputil
feat = torch.from_numpy(synth.make_feature_map(1)).cuda()
q = torch.randn(32, 30, device='cuda')
for n in (30, 100):
    q = torch.randn(32, n, device='cuda')
    for _ in range(3): rputil.match_sample(q, feat, 2)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): rputil.match_sample(q, feat, 2)
    torch.cuda.synchronize(); print("match_sample n=%d: %.3f ms per call (incl. D2H of the points)" % (n, (time.perf_counter() - t) / 20 * 1e3))
for B in (2, 64):
    views = torch.from_numpy(np.concatenate([synth.make_warp_view(s % 4, 'suncg') for s in range(B)])).cuda()
    Rs = np.stack([synth.make_pose(s) for s in range(B)])
    for _ in range(3): util.warping_device(views, Rs, 'suncg')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = torch.empty_like(views); Rt = torch.from_numpy(Rs.reshape(B, 16)).cuda()
    t = time.perf_counter()
    for _ in range(20): util.warping_device(views, Rs, 'suncg', out=out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
    print("warp B=%d views: %.3f ms per call (%.1f GB/s of 65 B/pixel algorithmic)" % (B, dt * 1e3, B * 102400 * 65 / dt / 1e9))
