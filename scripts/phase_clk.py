"""Per-phase SM cycles of the fused solver for one scan pair (rp_debug.phase_clk), T = 128 and T = 512 builds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import _lib, synth
from relativepose_b200.RPModule.rputil import opts
from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts
lib = _lib.load()
pl = [params_from_opts(opts(*synth.shipped_params("suncg")[0]))]
sv = PoseSolver("cuda:0")
recs = synth.make_batch(7_000_000, 8, 103)
old = lib.rp_solver_wide_max(-1)
for B in (1, 8):
    d = PackedBatch(recs[:B]).to_device(sv.device)
    for name, wm in (("T=128", 0), ("T=512", 1 << 20)):
        lib.rp_solver_wide_max(wm)
        clk = torch.zeros((B, 8), dtype=torch.int64, device="cuda")
        dbg = _lib.RpDebug()
        dbg.phase_clk = clk.data_ptr()
        for _ in range(3):
            T, st, stats = sv.solve_device(d, pl, debug=dbg)
        torch.cuda.synchronize()
        c = clk.cpu().numpy().astype(np.float64)
        ph = np.stack([c[:, 1] - c[:, 0], c[:, 2] - c[:, 1], c[:, 3] - c[:, 2], c[:, 4] - c[:, 3], c[:, 5] - c[:, 4], c[:, 6], c[:, 5] - c[:, 4] - c[:, 6]], 1) / 1.965e3
        print("B=%d %s  us per phase (mean over pairs): A %.1f | B+C %.1f | D %.1f | E %.1f | F %.1f (eigen iterations %.1f, fits+rest %.1f) | total %.1f; power its %.0f"
              % ((B, name) + tuple(ph.mean(0)) + (float((c[:, 5] - c[:, 0]).mean() / 1.965e3), float(stats[:, 4].double().mean()))))
lib.rp_solver_wide_max(old)
