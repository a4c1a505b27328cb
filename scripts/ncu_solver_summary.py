"""profiles/r2_solver_bench_launch_ncu.txt + profiles/r2_solver_ncu.json (the constants bench.py's roofline uses) from an
`ncu --set full` capture of the bench launch.  usage: ncu_solver_summary.py <rep> <pairs per launch> [tag]"""
import csv, io, json, os, subprocess, sys
rep, pairs = sys.argv[1], int(sys.argv[2])
tag = sys.argv[3] if len(sys.argv) > 3 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
get = lambda k: (units[h.index(k)], vals[h.index(k)]) if k in h else ("", "nan")
out = ["%s, kernel %s: %d pairs per launch (bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline under" % (os.path.basename(rep), vals[h.index("Kernel Name")], pairs),
       "ncu --set full --clock-control none --import-source on -k regex:rp_solve_kernel -s 6 -c 1)", ""]
for k in want:
    u, v = get(k)
    out.append("  %-82s %-16s %s" % (k, u, v))
def f(k, scale=1.0):
    u, v = get(k)
    x = float(v.replace(",", ""))
    if u.lower().startswith("mbyte"): x *= 1e6
    if u.lower().startswith("gbyte"): x *= 1e9
    if u.lower().startswith("kbyte"): x *= 1e3
    return x * scale
J = {"source": "profiles/%s_solver_bench_launch_ncu.txt" % tag, "pairs": pairs, "dram_read_bytes": f("dram__bytes_read.sum"),
     "dram_write_bytes": f("dram__bytes_write.sum"), "warp_inst": f("smsp__inst_executed.sum"),
     "thread_inst": f("smsp__thread_inst_executed.sum") if "smsp__thread_inst_executed.sum" in h else None,
     "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "pipe_fp64_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
     "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active")}
out += ["", "per pair: %.0f warp-instructions, %s thread-instructions, %.0f DRAM bytes" % (
    J["warp_inst"] / pairs, ("%.2f M" % (J["thread_inst"] / pairs / 1e6)) if J["thread_inst"] else "n/a",
    (J["dram_read_bytes"] + J["dram_write_bytes"]) / pairs)]
open(os.path.join(ROOT, "profiles", "%s_solver_bench_launch_ncu.txt" % tag), "w").write("\n".join(out) + "\n")
json.dump(J, open(os.path.join(ROOT, "profiles", "%s_solver_ncu.json" % tag), "w"), indent=1)
print("\n".join(out))
