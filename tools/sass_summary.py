"""SASS evidence of the tcgen05 / TMEM / bulk-copy path: per kernel of the built library, counts of the mnemonics
B200_PROFILING.md lists (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP / UTMALDG = bulk copies, SYNCS =
mbarrier ops) plus registers from the cubin.  Writes profiles/sass_<tag>_summary.txt.

    python tools/sass_summary.py r02
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = glob.glob(os.path.join(ROOT, "plspm-python_b200", "plspm_b200", "libplspm_b200.so"))[0]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+?):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
KEYS = ("UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "IMAD.WIDE", "DFMA", "HMMA", "IMMA")
counts, fn = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and fn:
        op = m.group(1)
        counts[fn]["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[fn][k] += 1
out = ["SASS mnemonic counts per kernel of libplspm_b200.so (cuobjdump -sass; sm_100a).  UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,",
       "UBLKCP = cp.async.bulk (linear), UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier.  Kernels without any of them listed by name only.", ""]
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
for (fn, c), dn in zip(counts.items(), demangle):
    r = regs.get(fn, ("?", "?"))
    hits = ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k])
    out.append("%-70s  %5d instr  regs %s  static smem %s%s" % (dn.split("(")[0][:70], c["_n"], r[0], r[1], "  | " + hits if hits else ""))
path = os.path.join(ROOT, "profiles", "sass_%s_summary.txt" % tag)
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
