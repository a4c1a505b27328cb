"""Top SASS instructions (by warp-level executions) of a kernel in an .ncu-rep, with source lines."""
import csv, io, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines
INNER = "--inner" in sys.argv
if INNER: sys.argv.remove("--inner")
rep, so = sys.argv[1], sys.argv[2]
ksub = sys.argv[3] if len(sys.argv) > 3 else "rp_solve_kernel"
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 10**9)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi_ + 1:] if len(r) == len(hdr)]
base = int(body[0][0], 16)
lines = sass_lines(so, ksub)
tot = sum(float(r[col["Instructions Executed"]] or 0) for r in body)
sel = []
for r in body:
    off = int(r[0], 16) - base
    chain, sass = lines.get(off, ([("?", 0)], r[1]))
    outer = [l for f, l in chain if f.endswith(".cu")]
    ol = (outer[0] if INNER else outer[-1]) if outer else 0
    if lo <= ol <= hi:
        sel.append((float(r[col["Instructions Executed"]] or 0), off, sass, [l for f, l in chain if f.endswith('.cu')]))
print("selected %.3g of %.3g warp-instr" % (sum(s[0] for s in sel), tot))
# opcode histogram
from collections import Counter
c = Counter()
for n, off, sass, ch in sel:
    c[sass.split()[0] if not sass.startswith('@') else sass.split()[1]] += n
for op, n in c.most_common(25):
    print("%-18s %6.2f%%" % (op, 100 * n / tot))
