"""Throughput of the fused solver vs the number of resident CTAs (= workspace slots): fewer slots keep the per-slot working
set (~225 KB) inside the 126 MB L2, more slots hide more latency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import synth
from relativepose_b200.RPModule.rputil import opts
from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts
para = opts(*synth.shipped_params("suncg")[0])
recs = synth.make_batch(5_000_000, 256, 103)
recs = [recs[i % 256] for i in range(4096)]
pk = PackedBatch(recs)
for slots in (296, 444, 518, 592, 740):
    sv = PoseSolver("cuda:0", n_slots=slots)
    d = pk.to_device(sv.device)
    pl = [params_from_opts(para)]
    try:
        for _ in range(3): sv.solve_device(d, pl)
    except Exception as e:
        print(slots, "failed:", str(e)[:80]); continue
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): sv.solve_device(d, pl)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print("n_slots %4d: %.3f ms per 4096 pairs = %.0f k pairs/s" % (slots, dt * 1e3, 4096 / dt / 1e3))
