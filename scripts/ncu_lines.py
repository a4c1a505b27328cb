"""Per-source-line instruction / stall profile from an .ncu-rep (run where ncu + nvdisasm exist, no GPU needed).

  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep relativepose_b200/librp_b200.so [kernel-substring] [--top 40]

Joins `ncu --page source --csv` (SASS rows) with `nvdisasm -gi` line info of the cubin embedded in the .so,
attributing inlined code to the outermost line of the given source file.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
        cur_fn, cur = None, None
        prev_was_file = False
        for line in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
            if m:
                cur_fn = m.group(1)
                out.setdefault(cur_fn, {})
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
            if m:
                chain = [(m.group(1), int(m.group(2)))]
                for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
                    chain.append((mm.group(1), int(mm.group(2))))
                # consecutive "//## File" lines spell one inline chain, innermost first
                if prev_was_file and cur:
                    for fr in chain:
                        if cur[-1] != fr:
                            cur.append(fr)
                else:
                    cur = chain
                prev_was_file = True
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and cur_fn is not None:
                out[cur_fn][int(m.group(1), 16)] = (list(cur) if cur else [("?", 0)], m.group(2).strip())
                prev_was_file = False
    for fn, d in out.items():
        if kernel_sub in fn:
            return d
    raise SystemExit("kernel %s not found in %s" % (kernel_sub, list(out)))


def main():
    rep, so = sys.argv[1], sys.argv[2]
    ksub = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "rp_solve_kernel"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    srcfile = None
    sel = ["--launch-skip", sys.argv[sys.argv.index("--skip") + 1], "--launch-count", "1"] if "--skip" in sys.argv else []
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))   # first kernel only
    body = [r for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]
    base = int(body[0][0], 16)
    lines = sass_lines(so, ksub)
    agg = defaultdict(lambda: defaultdict(float))
    tot = defaultdict(float)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in body:
        off = int(r[0], 16) - base
        chain, sass = lines.get(off, ([("?", 0)], r[1]))
        # outermost frame in the main source file
        key = None
        for f, l in (chain if "--inner" in sys.argv else reversed(chain)):
            if f.endswith(".cu"):
                key = (os.path.basename(f), l)
                break
        if key is None:
            key = (os.path.basename(chain[-1][0]), chain[-1][1])
        n = float(r[col["Instructions Executed"]] or 0)
        s = float(r[col["# Samples"]] or 0)
        agg[key]["inst"] += n
        agg[key]["samples"] += s
        tot["inst"] += n
        tot["samples"] += s
        for h in stall_cols:
            v = float(r[col[h]] or 0)
            agg[key][h] += v
    src = {}
    for f in set(k[0] for k in agg):
        for cand in (os.path.join("relativepose_b200", "csrc", f),):
            if os.path.exists(cand):
                src[f] = open(cand).read().splitlines()
    print("total warp instructions %.3g, samples %.0f" % (tot["inst"], tot["samples"]))
    print("%-22s %8s %7s %7s  %-28s %s" % ("line", "inst%", "smpl%", "cum%", "top stalls", "source"))
    cum = 0.0
    for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        cum += d["samples"]
        st = sorted(((d[h], h[6:]) for h in stall_cols), reverse=True)[:2]
        sts = " ".join("%s:%.0f%%" % (n, 100 * v / max(d["samples"], 1)) for v, n in st)
        code = src.get(key[0], [""] * (key[1] + 1))
        text = code[key[1] - 1].strip()[:90] if 0 < key[1] <= len(code) else ""
        print("%-22s %7.2f%% %6.2f%% %6.1f%%  %-28s %s" % ("%s:%d" % key, 100 * d["inst"] / tot["inst"],
                                                             100 * d["samples"] / tot["samples"], 100 * cum / tot["samples"], sts, text))


if __name__ == "__main__":
    main()
