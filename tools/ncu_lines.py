"""Stall samples of one kernel launch aggregated by source line: joins `ncu --page source --print-source=sass --csv`
with the line table of the library's cubin (nvdisasm -g), instruction by instruction.

    cuobjdump -xelf all <lib.so>; nvdisasm -g -c <cubin> > all.sass
    ncu -i rep --page source --csv --print-source=sass -k regex:<kernel> --launch-skip S --launch-count 1 > k.csv
    python tools/ncu_lines.py all.sass <mangled kernel name substring> k.csv [top]
"""
import csv
import re
import sys
from collections import defaultdict

sass, kname, kcsv = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines, cur, inside = [], ("?", 0), False
for ln in open(sass):
    if ln.startswith(".text."):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
rows = list(csv.reader(open(kcsv)))
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # which launch in the file (ncu -i prints every matching launch)
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = hdrs[which]
h = rows[hdr]
si, ei = h.index("# Samples"), h.index("Instructions Executed")
body = [r for r in rows[hdr + 1:(hdrs[which + 1] if which + 1 < len(hdrs) else len(rows))] if len(r) > ei and r[0].startswith("0x")]
print("instructions: sass %d, ncu %d" % (len(lines), len(body)))
agg, ex, tot = defaultdict(int), defaultdict(int), 0
for k, r in enumerate(body):
    key = lines[k] if k < len(lines) else ("?", 0)
    s = int(r[si] or 0)
    agg[key] += s
    ex[key] += int(r[ei] or 0)
    tot += s
for key, s in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print("%6.2f%%  %8d samples  %10d warp-instr  %s:%d" % (100.0 * s / max(tot, 1), s, ex[key], key[0], key[1]))
