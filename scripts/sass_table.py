"""profiles/r2_sass_counts.txt: per-kernel counts of the SASS mnemonics that identify the Blackwell paths
(cuobjdump -sass of the shipped library; no GPU needed)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "relativepose_b200", "librp_b200.so")
MN = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FADD2", "FMUL2", "FFMA2",
      "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "ATOMS", "BAR"]
txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
cur, tab = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        cur = name.split("(")[0][-70:]
        tab[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        tab[cur]["total"] += 1
        for k in MN:
            if op == k or op.startswith(k + ".") or (k in ("LDS", "STS", "LDG", "STG", "BAR", "MUFU", "ATOMS") and op.startswith(k)):
                tab[cur][k] += 1
out = ["SASS mnemonic counts per kernel, cuobjdump -sass relativepose_b200/librp_b200.so (sm_100a)",
       "UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (1-D TMA), UTMALDG/UTMASTG = tiled TMA, UTCBAR = tcgen05.commit,",
       "SYNCS = mbarrier, FADD2/FMUL2 = packed f32x2 (solver front end), HMMA = legacy mma.sync (none expected)", ""]
cols = ["total"] + MN
out.append("%-72s" % "kernel" + "".join("%9s" % c for c in cols))
for k, c in tab.items():
    if c["total"] < 40:
        continue
    out.append("%-72s" % k + "".join("%9d" % c[x] for x in cols))
open(os.path.join(ROOT, "profiles", "r2_sass_counts.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:4] + [l[:200] for l in out[4:12]]))
