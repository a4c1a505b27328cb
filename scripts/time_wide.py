"""T = 128 build vs T = 512 build of the solver (rp_solver_wide_max) over batch sizes: time per batch, per stage for a single
pair, and the distance between the two builds' poses."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import _lib, synth
from relativepose_b200.RPModule.rputil import opts
from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts
lib = _lib.load()
para = opts(*synth.shipped_params("suncg")[0])
pl = [params_from_opts(para)]
sv = PoseSolver("cuda:0")
recs_all = synth.make_batch(7_000_000, 1184, 103)


def timed(d, n=20, **kw):
    for _ in range(3): sv.solve_device(d, pl, **kw)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = sv.solve_device(d, pl, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


default_max = lib.rp_solver_wide_max(-1)
print("default wide_max", default_max)
for B in (1, 4, 16, 32, 74, 148, 222, 296, 444, 592, 1184):
    d = PackedBatch(recs_all[:B]).to_device(sv.device)
    lib.rp_solver_wide_max(0)
    tn, (Tn, stn, _) = timed(d)
    lib.rp_solver_wide_max(1 << 30)
    tw, (Tw, stw, _) = timed(d)
    diff = float((Tn - Tw).abs().max())
    print("B %5d: T=128 %.3f ms   T=512 %.3f ms   ratio %.2f   max|dT| %.2e  status equal %s" % (B, tn, tw, tn / tw, diff, bool((stn == stw).all())))
d = PackedBatch(recs_all[:1]).to_device(sv.device)
for stage, nm in ((_lib.STAGE_TOPK, 'A'), (_lib.STAGE_AFFINITY, 'A-D'), (_lib.STAGE_SOLVE, 'A-F')):
    lib.rp_solver_wide_max(0); tn, _ = timed(d, stop_after=stage)
    lib.rp_solver_wide_max(1 << 30); tw, _ = timed(d, stop_after=stage)
    print("single pair, stop after %-4s: T=128 %.3f ms  T=512 %.3f ms" % (nm, tn, tw))
lib.rp_solver_wide_max(default_max)
