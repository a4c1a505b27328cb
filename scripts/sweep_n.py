"""configs[4]: N sweep {128..2048}: pairs/s, dense-equivalent GB/s and measured structure per N."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import synth
from relativepose_b200.RPModule.rputil import opts
from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts
para = opts(*synth.shipped_params("suncg")[0]); pl = [params_from_opts(para)]
rows = []
for nominal, pairs in ((128, 16384), (256, 8192), (512, 4096), (1024, 1184), (2048, 296)):
    n = synth.keypoints_for_nominal_N(nominal)
    uniq = min(pairs, 64)
    recs = synth.make_batch(6_000_000 + nominal, uniq, n)
    pk = PackedBatch([recs[i % uniq] for i in range(pairs)])
    sv = PoseSolver("cuda:0")
    d = pk.to_device(sv.device)
    for _ in range(2):
        T, st, stats = sv.solve_device(d, pl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        T, st, stats = sv.solve_device(d, pl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    s = stats.cpu().numpy()
    N = n * 5
    its = s[:, 4].mean()
    balg = 4.0 * N * N * (2 + its + 5) + 156.0 * 2 * n
    rows.append(dict(nominal_N=nominal, N_actual=N, pairs=pairs, ms=ms, pairs_per_s=pairs / ms * 1e3, survivors=float(s[:, 2].mean()),
                     power_its=float(its), dense_equiv_GBps=balg * pairs / (ms * 1e-3) / 1e9))
    print(json.dumps(rows[-1]))
    del sv, d
    torch.cuda.empty_cache()
