"""Run the fused solver a few times on a synthetic batch (for ncu / quick timing)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from relativepose_b200 import synth
from relativepose_b200.RPModule.rputil import opts
from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=592)
ap.add_argument("--n", type=int, default=103)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--method", default="irls+sm")
a = ap.parse_args()
para = opts(*synth.shipped_params("suncg")[0])
para.method = a.method
uniq = min(a.pairs, 128)
recs = synth.make_batch(5_000_000, uniq, a.n)
recs = [recs[i % uniq] for i in range(a.pairs)]
pk = PackedBatch(recs)
sv = PoseSolver("cuda:0")
d = pk.to_device(sv.device)
pl = [params_from_opts(para)]
for _ in range(a.iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    T, st, stats = sv.solve_device(d, pl)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pairs=%d n=%d  %.3f ms  %.0f pairs/s" % (a.pairs, a.n, dt * 1e3, a.pairs / dt))
s = stats.cpu().numpy()
print("stats mean: N=%.0f M1=%.0f M2=%.0f nz=%.0f its=%.1f maxit=%d notconv=%d" % (
    s[:, 0].mean(), s[:, 1].mean(), s[:, 2].mean(), s[:, 3].mean(), s[:, 4].mean(), s[:, 5].max(), s[:, 6].sum()))
