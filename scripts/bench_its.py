"""Distribution of power-iteration counts on the headline batch + throughput under RP_PI_FAST_CAP / RP_PI_SWITCH."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import synth
from relativepose_b200.RPModule.rputil import opts
from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts
pl = [params_from_opts(opts(*synth.shipped_params("suncg")[0]))]
sv = PoseSolver("cuda:0")
d = PackedBatch(synth.make_batch(1_000_000, 4096, 103)).to_device(sv.device)
for _ in range(3): T, st, stats = sv.solve_device(d, pl)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): T, st, stats = sv.solve_device(d, pl)
e1.record(); torch.cuda.synchronize()
s = stats.cpu().numpy()
mx = s[:, 5]
print("cap %s switch %s: %.3f ms per 4096 pairs (%.0f k pairs/s); max its per alternation: mean %.1f p99 %d max %d; robust pairs %d"
      % (os.environ.get("RP_PI_FAST_CAP", "-"), os.environ.get("RP_PI_SWITCH", "-"), e0.elapsed_time(e1) / 10, 4096 / (e0.elapsed_time(e1) / 10) , mx.mean(),
         int(np.percentile(mx, 99)), int(mx.max()), int(((s[:, 7] >> 8) & 1).sum())))
