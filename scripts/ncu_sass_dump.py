"""Dump SASS (address order) with executed counts for instructions whose innermost .cu line is in [lo,hi]."""
import csv, io, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines
rep, so, ksub, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi_ + 1:] if len(r) == len(hdr)]
base = int(body[0][0], 16)
lines = sass_lines(so, ksub)
for r in body:
    off = int(r[0], 16) - base
    chain, sass = lines.get(off, ([("?", 0)], r[1]))
    cu = [l for f, l in chain if f.endswith(".cu")]
    if cu and lo <= cu[0] <= hi:
        print("%06x %10s %6s  %-60s %s" % (off, r[col["Instructions Executed"]], r[col["# Samples"]], sass[:60], cu))
