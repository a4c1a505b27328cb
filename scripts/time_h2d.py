"""Pinned host -> HBM rate on this box and the e2e solve time against the number of pipeline chunks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relativepose_b200 import synth
from relativepose_b200.RPModule.rputil import opts
from relativepose_b200.solver import PackedBatch, PoseSolver
dev = torch.device("cuda:0")
for mb in (8, 64, 155):
    h = torch.empty((mb << 20,), dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D %4d MB pinned: %.1f GB/s" % (mb, mb * 1.048576e-3 * 10 / (e0.elapsed_time(e1) * 1e-3)))
para = opts(*synth.shipped_params("suncg")[0])
recs = synth.make_batch(1_000_000, 4096, 103)
pk = PackedBatch(recs)
sv = PoseSolver(dev)
for chunks in (1, 2, 4, 6, 7, 8, 12, 16):
    for _ in range(3): sv.solve_packed(pk, para, chunks=chunks)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): sv.solve_packed(pk, para, chunks=chunks)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print("chunks %2d: %.3f ms per 4096 pairs = %.0f k pairs/s e2e" % (chunks, dt * 1e3, 4096 / dt / 1e3))
