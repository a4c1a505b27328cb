"""Instruction/sample share of source regions, attributing by the INNERMOST .cu frame (function-level view)
and by the outermost (phase-level view).  usage: ncu_regions.py rep so kernel name:lo-hi ..."""
import csv, io, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines
rep, so, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
regs = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regs.append((n, int(lo), int(hi)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi_ + 1:] if len(r) == len(hdr)]
base = int(body[0][0], 16)
lines = sass_lines(so, ksub)
tot_i = tot_s = 0.0
inner = {n: [0.0, 0.0] for n, _, _ in regs}; outer = {n: [0.0, 0.0] for n, _, _ in regs}
for r in body:
    off = int(r[0], 16) - base
    chain, sass = lines.get(off, ([("?", 0)], r[1]))
    cu = [l for f, l in chain if f.endswith(".cu")]
    n_i = float(r[col["Instructions Executed"]] or 0); n_s = float(r[col["# Samples"]] or 0)
    tot_i += n_i; tot_s += n_s
    if not cu:
        continue
    for n, lo, hi in regs:
        if lo <= cu[0] <= hi: inner[n][0] += n_i; inner[n][1] += n_s
        if lo <= cu[-1] <= hi: outer[n][0] += n_i; outer[n][1] += n_s
print("total %.4g warp-instr, %d samples" % (tot_i, tot_s))
print("%-18s %10s %8s | %10s %8s" % ("region", "inner inst%", "smpl%", "outer inst%", "smpl%"))
for n, lo, hi in regs:
    print("%-18s %9.2f%% %7.2f%% | %9.2f%% %7.2f%%" % (n, 100 * inner[n][0] / tot_i, 100 * inner[n][1] / tot_s,
                                                       100 * outer[n][0] / tot_i, 100 * outer[n][1] / tot_s))
