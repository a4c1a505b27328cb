#!/usr/bin/env python
"""bench.py -- scan-pairs/sec of the RPModule hot path (BASELINE.json metric) + the other BASELINE configs as extras.

Headline workload (configs[1]): SUNCG-shape pair, nominal N=512 candidate matches (n_s=n_t=103, topK=5 ->
N_actual=515), RPModule only, default method irls+sm, sigmas = row 0 of the shipped SUNCG parameter file.  A *step* =
one fused launch over a batch of `--pairs` DISTINCT seeded scan pairs per GPU (weak scaling: every rank gets its own
batch; no collective on the data path).

  python bench.py [--gpus N --steps K --warmup W]          our arm (CUDA)
  python bench.py --impl reference [...]                    the reference's own CPU code (baseline/_ref or /root/reference
                                                            through oracle/ref_loader.py; the numpy port only if absent)

Prints ONE JSON line (rank 0).  Beside the contract keys it carries `extra`: driver-run numbers for the other configs
(SCNet / Resnet18_8s forward with tensor-core rooflines, configs[2], configs[3] 1- and 3-step, the ragged batch, the N
sweep of configs[4], and `c4` = 256 ScanNet-shape pairs through the 3-step alternation, sharded over the ranks).
See DESIGN.md section "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import time

# one BLAS/OpenMP thread per process: the CPU arm runs one worker process per core (must be set before numpy loads)
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scan-pairs/sec (R,t solved) at N=512 corr"
UNIT = "pairs/s"
N_NOMINAL = 512
TOPK = 5
SCNET_GFLOP_PER_PAIR = 72.27      # 2 x 36.14 GMAC per pair per alternation step
RESNET_GFLOP_PER_PAIR = 16.13     # Resnet18_8s, two images

# ncu --set full of the bench launch (profiles/r2_solver_bench_launch_ncu.txt); per 4096-pair launch of rp_solve_kernel<0>
NCU_SOLVER = {"source": "profiles/r2_solver_bench_launch_ncu.txt", "pairs": 4096,
              "dram_read_bytes": 395.87e6, "dram_write_bytes": 828.95e6, "warp_inst": 3.3204e9,
              "issue_active_pct": 42.5, "pipe_fp64_pct": 14.7, "warps_active_pct": 24.4}
_ncu_json = os.path.join(ROOT, "profiles", "r2_solver_ncu.json")
if os.path.exists(_ncu_json):
    NCU_SOLVER.update(json.load(open(_ncu_json)))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--pairs", type=int, default=4096, help="scan pairs per GPU per step")
    ap.add_argument("--nominal-n", type=int, default=N_NOMINAL)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU-baseline sample (0 = sized for ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra configs (SCNet / ResNet / configs[2-4] / c4)")
    ap.add_argument("--c4-pairs", type=int, default=256, help="total ScanNet-shape pairs of the sharded configs[3] run")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------------------------- CPU arm
_ref_state = {}


def _reference_solver():
    """('reference', callable) when the reference's own code is importable here (oracle/ref_loader.py finds /root/reference
    or baseline/_ref), else ('port', callable) -- the numpy restatement proven equal to it (tests/test_oracle_golden.py)."""
    if not _ref_state:
        from oracle import ref_loader, rp_oracle
        try:
            if not ref_loader.reference_available():
                raise FileNotFoundError(ref_loader.REFERENCE_ROOT)
            mod = ref_loader.load_reference_rpmodule()
            _ref_state["kind"] = "reference"
            _ref_state["fn"] = lambda s, t, sig: mod.RelativePoseEstimation_helper(s, t, ref_loader.reference_opts(*sig))
            _ref_state["where"] = ref_loader.REFERENCE_ROOT
        except Exception as e:                                  # noqa: BLE001  (absent tree, missing dependency of the tree)
            _ref_state["kind"] = "port"
            _ref_state["fn"] = lambda s, t, sig: rp_oracle.solve_pair(s, t, rp_oracle.Params(*sig))
            _ref_state["where"] = "oracle/rp_oracle.py (%s: %s)" % (type(e).__name__, e)
    return _ref_state


def _cpu_worker(job):
    from relativepose_b200 import synth
    st = _reference_solver()
    seed, n, sig = job
    rec = synth.make_pair(seed, n)
    s, t = synth.record_to_dicts(rec)
    t0 = time.perf_counter()
    T = st["fn"](s, t, sig)
    return time.perf_counter() - t0, T, st["kind"], st["where"]


def cpu_baseline(n_kp, sample, first_seed=10_000_000, target_s=15.0):
    """The reference's RelativePoseEstimation_helper (rpmodule.py:317-508) on all host cores, one process per core (the
    reference's own way to use more than one core: --entrySplit process sharding, evaluation.py:59).
    sample <= 0: sized from the warm-up pass for about `target_s` seconds of wall time (bounded sample of the workload)."""
    import logging
    import multiprocessing as mp
    from relativepose_b200 import synth
    logging.disable(logging.INFO)
    cores = os.cpu_count() or 1
    sig = tuple(float(x) for x in synth.shipped_params("suncg")[0])
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        warm = pool.map(_cpu_worker, [(first_seed - 1 - i, n_kp, sig) for i in range(cores)])   # imports, BLAS init
        if sample <= 0:
            lat0 = float(np.median([r[0] for r in warm]))
            sample = int(min(max(4 * cores, cores * target_s / max(lat0, 1e-4)), 20000))
        jobs = [(first_seed + i, n_kp, sig) for i in range(sample)]
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=max(1, sample // (cores * 16)))
        wall = time.perf_counter() - t0
    lat = np.array([r[0] for r in res])
    return {"value": sample / wall, "unit": UNIT, "cores": cores, "kind": res[0][2], "code": res[0][3],
            "sample": "%d pairs n_s=n_t=%d, one process per core, %.1f s wall" % (sample, n_kp, wall),
            "p50_ms_single_core": float(np.median(lat) * 1e3)}, wall, sample


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from relativepose_b200 import synth
    n_kp = synth.keypoints_for_nominal_N(args.nominal_n, TOPK)
    walls, counts = [], []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, wall, cnt = cpu_baseline(n_kp, args.cpu_sample, first_seed=20_000_000 + 100_000 * i, target_s=8.0)
        if i >= args.warmup:
            walls.append(wall)
            counts.append(cnt)
    value = sum(counts) / sum(walls)
    cb["value"] = value
    cfg = workload_config(args, n_kp, int(round(np.mean(counts))))
    cfg["pairs_per_step"] = "a bounded sample per step (mean %d pairs, ~8 s of CPU work on %d cores); the CUDA arm's step is %d " \
                            "pairs of the same generator" % (int(round(np.mean(counts))), cb["cores"], args.pairs)
    cfg["l2"] = "n/a (CPU)"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": cfg, "cpu_baseline": cb,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------- helpers
def workload_config(args, n_kp, pairs):
    return {"workload": "configs[1]: SUNCG-shape pair, RPModule only (irls+sm), nominal N=%d -> n_s=n_t=%d, topK=%d, "
                        "N_actual=%d; params = final_param_suncg_rlevel_3.txt row 0" % (args.nominal_n, n_kp, TOPK, n_kp * TOPK),
            "pairs_per_gpu_per_step": pairs, "distinct_pairs": pairs,
            "l2": "inputs exceed the 126 MB L2 (inputs %.0f MB/step, all pairs distinct)" % (pairs * (n_kp * 2 * (7 * 8 + 32 * 4)) / 1e6),
            "parallelism": "pair-sharded, no collective"}


class ClockSampler(object):
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.p = None
        self.index = index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self, window=None):
        """window = (t0, t1) wall-clock bounds (time.time()) of the timed region: only samples inside it are used (the
        sampler is started well before, nvidia-smi needs ~0.2 s to emit its first line); falls back to all samples."""
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if window and window[0] - 0.02 <= r[0] <= window[1] + 0.02]
        use = inside if inside else rows
        sm, mx, reasons = [r[1] for r in use], [r[2] for r in use], set()
        for r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0,
            "source": "fallback (B200_PROFILING.md)"}


def device_time_ms(torch, fn, n, warm):
    """Mean CUDA-event time of fn() on the current stream."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def wall_ms(torch, fn, n, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


def wall_ms_median(torch, fn, n, warm):
    """Median wall time of n individually timed calls (host-synchronous pipelines: one slow call -- an allocator refill, a lazy
    capture -- must not pass for the steady state); also returns the slowest."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts)), float(max(ts))


# --------------------------------------------------------------------------------------------- extras (rank 0, one GPU)
def synth_scans(B, seed=0):
    """B scan pairs (2B images) of the shape SURVEY 8(d) describes: rgb U(0,1), unit normals, smooth depth."""
    rs = np.random.RandomState(seed)
    K = 103
    pts = np.stack((rs.uniform(1, 637, (2 * B, K)), rs.uniform(1, 157, (2 * B, K))), 2)
    w = np.where((pts[..., 0] >= 160) & (pts[..., 0] <= 320), 1.0, 0.99)
    nrm = rs.randn(2 * B, 160, 640, 3).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=3, keepdims=True)
    yy, xx = np.mgrid[0:160, 0:640]
    depth = np.stack([2.5 + 1.5 * np.sin(xx / 37.0 + i) * np.cos(yy / 23.0) for i in range(2 * B)]).astype(np.float32)
    rgb = rs.uniform(0, 1, (2 * B, 160, 640, 3)).astype(np.float32)
    return rgb, nrm, depth, pts, w


def run_extras(torch, dev, peaks, steps):
    import types
    from relativepose_b200 import pipeline, synth
    from relativepose_b200.model.mymodel import SCNet, Resnet18_8s
    from relativepose_b200.RPModule.rputil import opts
    from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts
    ex = {}
    sus = peaks["bf16_tflops_sustained"]
    n = max(3, min(steps, 10))
    B = 32
    rgb, nrm, depth, pts, w = synth_scans(B)

    # ---- M1: SCNet forward, 32 scan pairs (SUNCG head layout), default bf16 tcgen05 path, CUDA-graph replay
    torch.manual_seed(0)
    cnet = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=1, skipLayer=1, outputType='rgbdnsf', snumclass=15)).to(dev)
    x16 = torch.cat([torch.from_numpy(synth.make_panorama_pair(s, "suncg")) for s in range(B)], 0).to(dev)
    ms = device_time_ms(torch, lambda: cnet(x16), n, 5)
    tf = SCNET_GFLOP_PER_PAIR * B / ms
    ex["scnet"] = {"workload": "SCNet.forward, %d scan pairs [64,16,160,640] -> [64,54,160,640], bf16 tcgen05 (fp32 accumulate)" % B,
                   "ms": ms, "pair_steps_per_s": B / ms * 1e3, "tflops": tf, "peak_tflops": sus, "frac": tf / sus,
                   "peak_source": peaks["source"] + " bf16_tflops_sustained", "gflop_per_pair_step": SCNET_GFLOP_PER_PAIR,
                   "ncu": "profiles/r2_scnet_halo_ncu_summary.txt (sm__pipe_tensor_cycles_active per layer)"}
    # the same forward in the split-precision tensor-core mode (three tcgen05 launches per layer, float32 storage): the parity mode
    from relativepose_b200.scnet_engine import ScnetEngine
    eng3 = ScnetEngine(cnet, mode='tc3')
    ms3 = device_time_ms(torch, lambda: eng3.forward(x16, borrow=True), n, 5)
    ex["scnet_tc3"] = {"workload": "SCNet.forward, %d scan pairs, RP_SCNET_MODE=tc3: half(x) w_hi + lo(x) w_hi + half(x) lo(w) on tcgen05, "
                                   "float32 storage (descriptors within 2.2e-4 of the reference's, poses within 2.3e-4 per step)" % B,
                       "ms": ms3, "pair_steps_per_s": B / ms3 * 1e3, "useful_tflops": SCNET_GFLOP_PER_PAIR * B / ms3,
                       "issued_tflops": 3 * SCNET_GFLOP_PER_PAIR * B / ms3}
    del x16, eng3

    # ---- M2: Resnet18_8s forward, 64 images
    torch.manual_seed(0)
    fnet = Resnet18_8s(types.SimpleNamespace(num_input=7, useTanh=1)).to(dev)
    x7 = torch.cat((torch.from_numpy(rgb), torch.from_numpy(nrm), torch.from_numpy(depth).unsqueeze(3)), 3).permute(0, 3, 1, 2).contiguous().to(dev)
    ms = device_time_ms(torch, lambda: fnet(x7), n, 5)
    tf = RESNET_GFLOP_PER_PAIR * B / ms
    ex["resnet18_8s"] = {"workload": "Resnet18_8s.forward, 64 images [64,7,160,640] -> [64,32,160,640], bf16 tcgen05", "ms": ms,
                         "tflops": tf, "peak_tflops": sus, "frac": tf / sus, "gflop_per_pair": RESNET_GFLOP_PER_PAIR}

    # ---- configs[2]: 32 Matterport-shape pairs, feature net + RPModule
    dep_d, nrm_d = torch.from_numpy(depth).to(dev), torch.from_numpy(nrm).to(dev)
    pts_d, w_d = torch.from_numpy(pts).to(dev), torch.from_numpy(w).to(dev)
    para = opts(*synth.shipped_params('matterport')[0])
    para.sigmaFeat = 0.05            # untrained descriptors: widen the soft-match kernel so rows do not all underflow

    def cfg2():
        return pipeline.solve_from_maps(fnet(x7), dep_d, nrm_d, pts_d, w_d, para, 'matterport')
    ms = wall_ms(torch, cfg2, n, 3)
    f_ = fnet(x7)
    ms_solve = wall_ms(torch, lambda: pipeline.solve_from_maps(f_, dep_d, nrm_d, pts_d, w_d, para, 'matterport'), n, 3)
    ex["config2"] = {"workload": "configs[2]: 32 Matterport-shape pairs, Resnet18_8s + gather + RPModule (N=515), device-resident "
                                 "maps, poses to the host", "ms": ms, "pairs_per_s": B / ms * 1e3, "ms_gather_solve_d2h": ms_solve}
    del fnet, x7, f_

    # ---- configs[3]: 32 ScanNet-shape pairs per GPU, SCNet + RPModule, 1 and 3 alternation steps (host scans in)
    torch.manual_seed(0)
    snet = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=1, outputType='rgbdnsf', snumclass=21)).to(dev)
    P = synth.shipped_params('scannet')
    for st_ in (1, 3):
        pa = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
        a = types.SimpleNamespace(snumclass=21, featureDim=32, outputType='rgbdnsf', maskMethod='kinect', alterStep=st_,
                                  dataset='scannet', para=pa, representation='skybox', completion=True)
        ms, ms_max = wall_ms_median(torch, lambda: pipeline.RelativePoseEstimationViaCompletion_batch(snet, rgb, nrm, depth, pts, w, a), 5, 4)
        ex["config3_%dstep" % st_] = {"workload": "configs[3] per GPU: 32 ScanNet-shape pairs, %d x (warp -> SCNet -> blend -> gather -> "
                                                  "RPModule N=515), host scans in, poses out (median of 5 calls)" % st_, "ms": ms,
                                      "ms_slowest_call": ms_max, "pairs_per_s": B / ms * 1e3}
    # the same 3-step alternation with the network in its split-precision (parity) mode
    from relativepose_b200.scnet_engine import ScnetEngine
    snet._engine = ScnetEngine(snet, mode='tc3')
    ms, ms_max = wall_ms_median(torch, lambda: pipeline.RelativePoseEstimationViaCompletion_batch(snet, rgb, nrm, depth, pts, w, a), 5, 4)
    ex["config3_3step_tc3"] = {"workload": "configs[3] per GPU, 3 steps, network in RP_SCNET_MODE=tc3 (split-precision tcgen05, the parity mode; "
                                           "median of 5 calls)", "ms": ms, "ms_slowest_call": ms_max, "pairs_per_s": B / ms * 1e3}
    del snet, cnet
    torch.cuda.empty_cache()

    # ---- ragged batch: 2048 distinct pairs, n_s, n_t ~ U{60..140} independently (load imbalance across the persistent CTAs)
    rs = np.random.RandomState(7)
    ns_l, nt_l = rs.randint(60, 141, 2048), rs.randint(60, 141, 2048)
    recs = [synth.make_pair(3_000_000 + i, int(ns_l[i]), int(nt_l[i])) for i in range(2048)]
    pk = PackedBatch(recs)
    sv = PoseSolver(dev)
    pl = [params_from_opts(opts(*synth.shipped_params("suncg")[0]))]
    d = pk.to_device(dev)
    ms = device_time_ms(torch, lambda: sv.solve_device(d, pl), n, 3)
    ex["ragged"] = {"workload": "2048 distinct pairs, n_s and n_t drawn independently from {60..140} (mean N = 500)", "ms": ms,
                    "pairs_per_s": 2048 / ms * 1e3}
    del sv, d, pk, recs

    # ---- configs[4]: N sweep
    rows = []
    for nominal, pairs in ((128, 16384), (256, 8192), (512, 4096), (1024, 1184), (2048, 296)):
        nk = synth.keypoints_for_nominal_N(nominal)
        uniq = min(pairs, 128)
        rr = synth.make_batch(6_000_000 + nominal, uniq, nk)
        pk = PackedBatch([rr[i % uniq] for i in range(pairs)])
        sv = PoseSolver(dev)
        d = pk.to_device(dev)
        ms = device_time_ms(torch, lambda: sv.solve_device(d, pl), 3, 2)
        _, _, stats = sv.solve_device(d, pl)
        s = stats.cpu().numpy()
        in_bytes = pairs * 2 * nk * (7 * 8 + 32 * 4)
        rows.append({"nominal_N": nominal, "N_actual": nk * 5, "pairs": pairs, "ms": ms, "pairs_per_s": pairs / ms * 1e3,
                     "surviving_pairs_mean": float(s[:, 2].mean()), "power_its_mean": float(s[:, 4].mean()),
                     "input_stream_GBps": in_bytes / (ms * 1e-3) / 1e9})
        del sv, d, pk
        torch.cuda.empty_cache()
    ex["sweep"] = {"workload": "configs[4]: N sweep, RPModule only, device-resident; input_stream_GBps = the bytes that must cross "
                               "HBM (inputs) / time -- the solver is issue-bound, not HBM-bound (see roofline)", "rows": rows}
    return ex


def run_c4(torch, dev, rank, world, total_pairs, barrier, max_over_ranks):
    """configs[3] as north_star names it: `total_pairs` ScanNet-shape pairs through the 3-step alternation (SCNet + RPModule),
    sharded over the ranks in contiguous blocks (strong scaling, no collective), 32 pairs per network call."""
    import types
    from relativepose_b200 import pipeline, sharding, synth
    from relativepose_b200.model.mymodel import SCNet
    from relativepose_b200.RPModule.rputil import opts
    lo, hi = sharding.shard_bounds(total_pairs, rank, world)
    mine = hi - lo
    torch.manual_seed(0)
    snet = SCNet(types.SimpleNamespace(batchnorm=1, useTanh=0, skipLayer=1, outputType='rgbdnsf', snumclass=21)).to(dev)
    P = synth.shipped_params('scannet')
    pa = opts(P[:3, 0], P[:3, 1], P[:3, 2], np.array([0.05, 0.05, 0.05]))
    a = types.SimpleNamespace(snumclass=21, featureDim=32, outputType='rgbdnsf', maskMethod='kinect', alterStep=3,
                              dataset='scannet', para=pa, representation='skybox', completion=True)
    CH = 32
    base = synth_scans(CH, seed=100 + rank)
    reps_ = -(-mine // CH)
    rgb, nrm, depth, pts, w = [np.concatenate([arr] * reps_, 0)[:2 * mine] if reps_ > 1 else arr[:2 * mine] for arr in base]

    def one_pass():
        # the rank's pairs in ONE call: the network runs CH pairs at a time, every alternation step solves all pairs at once
        return pipeline.RelativePoseEstimationViaCompletion_batch(snet, rgb, nrm, depth, pts, w, a, chunk=CH)
    for _ in range(2):
        one_pass()
    barrier()
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        one_pass()
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0) / reps
    return {"workload": "configs[3] (BASELINE 'batch 256 ScanNet-shape pairs, completion U-Net + RPModule, sharded'): %d pairs total, "
                        "3 alternation steps, contiguous blocks over %d rank(s), one call per rank (32 pairs per SCNet call, one solve per step over the rank's pairs), host scans in / poses out"
                        % (total_pairs, world), "scaling": "strong", "pairs_total": total_pairs, "pairs_this_rank": mine,
            "ms": wall * 1e3, "pairs_per_s": total_pairs / wall}


# --------------------------------------------------------------------------------------------- CUDA arm
def pin_to_gpu_numa_node(torch, local_rank):
    """Multi-GPU runs: keep this rank's threads (and therefore the pages of its pinned staging buffers, first touch) on the NUMA
    node its GPU hangs off -- eight ranks each streaming 155 MB per step out of host memory otherwise cross the socket link.
    Best effort: returns a description, or None when the topology is not exposed (single node, container without sysfs)."""
    try:
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bdf is None:
            import subprocess
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 stdout=subprocess.PIPE, text=True, timeout=10).stdout.strip()
            bdf = out.splitlines()[0].strip() if out else None
        if not bdf:
            return None
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:                    # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if len(cpus) < 2 or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None


def run_cuda_arm(args):
    import torch
    from relativepose_b200 import _lib, synth
    from relativepose_b200.RPModule.rputil import opts
    from relativepose_b200.solver import PackedBatch, PoseSolver, params_from_opts

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=cuda) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n_kp = synth.keypoints_for_nominal_N(args.nominal_n, TOPK)
    B = args.pairs
    para = opts(*synth.shipped_params("suncg")[0])
    plist = [params_from_opts(para)]

    # synthetic batch: B DISTINCT seeded pairs (distinct data per rank)
    recs = synth.make_batch(1_000_000 * (rank + 1), B, n_kp)
    packed = PackedBatch(recs)
    solver = PoseSolver(dev)
    dbatch = packed.to_device(dev)
    torch.cuda.synchronize()
    lib = _lib.load()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---------------- device-resident throughput ("value")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # before the warm-up: nvidia-smi needs a moment to emit its first sample
    for _ in range(args.warmup):
        solver.solve_device(dbatch, plist)
    barrier()
    if rank == 0:
        time.sleep(0.3)
    launches0 = lib.rp_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for i in range(args.steps):
        evs[i][0].record()
        T, status, stats = solver.solve_device(dbatch, plist)
        evs[i][1].record()
    barrier()
    wall = time.perf_counter() - t0
    w1 = time.time()
    launches = lib.rp_launch_count() - launches0
    clocks = sampler.stop((w0, w1)) if rank == 0 else None
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    wall = max_over_ranks(wall)
    value = world * B * args.steps / wall

    st = stats.cpu().numpy()
    status_h = status.cpu().numpy()
    kern_s = float(np.mean(kern_ms)) * 1e-3

    # ---------------- end to end through the C-ABI batch call with host buffers ("e2e": packed + pinned wire format)
    for _ in range(max(1, args.warmup)):
        solver.solve_packed(packed, para)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Th = solver.solve_packed(packed, para)
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * args.steps / e2e_wall

    # ---------------- end to end through the reference-named entry on a list of record dicts (packing + pinning timed)
    from relativepose_b200.RPModule.rpmodule import RelativePoseEstimation_batch, RelativePoseEstimation_helper
    RelativePoseEstimation_batch(recs, para)
    barrier()
    t0 = time.perf_counter()
    nrec = max(2, min(args.steps, 5))
    for _ in range(nrec):
        Tr = RelativePoseEstimation_batch(recs, para)
    barrier()
    rec_wall = max_over_ranks(time.perf_counter() - t0)
    e2e_records = world * B * nrec / rec_wall

    # ---------------- single-pair latency through the reference-named call
    s1, t1 = synth.record_to_dicts(recs[0])
    lat = []
    for i in range(40):
        torch.cuda.synchronize()
        a = time.perf_counter()
        RelativePoseEstimation_helper(s1, t1, para)
        lat.append(time.perf_counter() - a)
    p50 = float(np.median(lat[5:]) * 1e3)

    # ---------------- parity of the benchmarked batch against the oracle, on EVERY rank (4 pairs each; max over ranks)
    from oracle import rp_oracle
    Tdev = T.cpu().numpy()
    errs = []
    for b in (0, B // 3, 2 * B // 3, B - 1):
        s, t = synth.record_to_dicts(recs[b])
        To = rp_oracle.solve_pair(s, t, rp_oracle.Params(*synth.shipped_params("suncg")[0]))
        errs.append(float(max(np.linalg.norm(Tdev[b] - To), np.linalg.norm(Th[b] - To), np.linalg.norm(Tr[b] - To))))
    parity_err = max_over_ranks(max(errs))
    ok_frac = -max_over_ranks(-float((status_h == 0).mean()))          # min over ranks

    # ---------------- extras
    peaks = load_peaks()
    extra = {}
    if not args.no_extra:
        extra["c4"] = run_c4(torch, dev, rank, world, args.c4_pairs, barrier, max_over_ranks)
        if world == 1:
            extra.update(run_extras(torch, dev, peaks, args.steps))

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel (rp_solve_kernel<0>): instruction issue, with the HBM view beside it
    warp_inst = NCU_SOLVER["warp_inst"] / NCU_SOLVER["pairs"] * B
    issue_peak = 148 * 4 * peaks["sm_max_mhz"] * 1e6 / 1e9                  # G warp-instructions / s at the max SM clock
    issue_ach = warp_inst / kern_s / 1e9
    traffic = (NCU_SOLVER["dram_read_bytes"] + NCU_SOLVER["dram_write_bytes"]) / NCU_SOLVER["pairs"] * B
    in_bytes = packed.h2d_bytes() + B * (16 * 8 + 4)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, n_kp, B),
        "per_pair_p50_ms": p50,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": packed.h2d_bytes() + B * TOPK * 4,
                "d2h_bytes_per_step": B * (16 * 8 + 4), "api": "PoseSolver.solve_packed (C-ABI rp_solve_batch_ex, pinned host buffers)"},
        "e2e_records": {"value": e2e_records, "unit": UNIT, "api": "RPModule.rpmodule.RelativePoseEstimation_batch(list of record "
                        "dicts): concatenation + pinning + H2D + solve + D2H timed"},
        "gpu_launches": int(launches),
        "host_affinity": numa if numa else ("not pinned (single GPU)" if world == 1 else "not pinned (no NUMA topology exposed)"),
        "clocks": clocks,
        "roofline": {
            "bound": "issue", "kernel": "rp_solve_kernel<false,false>", "kernel_ms": float(np.mean(kern_ms)),
            "achieved": issue_ach, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": issue_ach / issue_peak,
            "peak_source": "148 SMs x 4 schedulers x 1 warp-instruction/clk x %.0f MHz (max SM clock, %s)" % (peaks["sm_max_mhz"], peaks["source"]),
            "model": "warp-instructions executed per pair (ncu smsp__inst_executed.sum / pairs, %s) x pairs per launch / CUDA-event "
                     "time of the launch; the affinity is a ~2 k-edge CSR per pair kept on chip, so the kernel is neither HBM- nor "
                     "tensor-bound" % NCU_SOLVER["source"],
            "warp_inst_per_pair": NCU_SOLVER["warp_inst"] / NCU_SOLVER["pairs"],
            "ncu_issue_active_pct": NCU_SOLVER["issue_active_pct"], "ncu_pipe_fp64_pct": NCU_SOLVER["pipe_fp64_pct"],
            "ncu_warps_active_pct": NCU_SOLVER["warps_active_pct"],
            "traffic": traffic, "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the bench launch (%s), per pair x "
                                                  "pairs per launch" % NCU_SOLVER["source"],
            "hbm": {"algorithmic_bytes": in_bytes, "algorithmic_GBps": in_bytes / kern_s / 1e9, "actual_GBps": traffic / kern_s / 1e9,
                    "peak_GBps": peaks["hbm_gbs"], "hbm_actual_frac": traffic / kern_s / 1e9 / peaks["hbm_gbs"],
                    "traffic_over_algorithmic": traffic / in_bytes, "peak_source": peaks["source"] + " hbm_gbs",
                    "note": "algorithmic bytes = the input stream + poses (what must cross HBM once)"},
            "mean_power_iters_per_pair": float(st[:, 4].astype(np.float64).mean())},
        "parity": {"status_ok_frac_min_over_ranks": ok_frac, "max_T_frobenius_err_vs_oracle": parity_err,
                   "checked": "4 pairs per rank x {device-resident, e2e, e2e_records} results, max over ranks"},
        "extra": extra,
    }
    if not args.no_cpu_baseline and world == 1:
        cb, _, _ = cpu_baseline(n_kp, args.cpu_sample)
        out["cpu_baseline"] = cb
    print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
