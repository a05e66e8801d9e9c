"""Attribute an ncu SASS-page CSV (instructions executed, stall samples) to source lines.

    cuobjdump -xelf all lib.so ; nvdisasm -g -c X.cubin > all.sass
    ncu -i rep.ncu-rep --page source --csv > sass.csv
    python tools/ncu_by_line.py all.sass '<mangled kernel name>' sass.csv [top]

The SASS page has no line column, so the instruction stream of the profiled kernel is aligned, in order, with
nvdisasm's line-annotated listing of the same cubin (the opcodes are cross-checked).
"""
import collections
import csv
import re
import sys

sass_path, kernel, csv_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by_samples = len(sys.argv) > 5 and sys.argv[5] == "samples"  # order the per-line table by stall samples

lines = open(sass_path).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":"))
listing = []  # (file, line, opcode)
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        parts = m.group(1).split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        listing.append((cur[0], cur[1], op))

rows = list(csv.reader(open(csv_path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
prof = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) > max(ci, cs):
        parts = r[ct].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        prof.append((op, int(r[ci] or 0), int(r[cs] or 0)))
if len(prof) != len(listing):
    print(f"warning: {len(prof)} profiled instructions vs {len(listing)} in the listing", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0])
mism = 0
for (f, ln, op), (pop, n, s) in zip(listing, prof):
    mism += op.split(".")[0] != pop.split(".")[0]
    agg[(f, ln)][0] += n
    agg[(f, ln)][1] += s
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print(f"opcode mismatches {mism}; total inst {ti}, samples {ts}")
byfile = collections.defaultdict(lambda: [0, 0])
for (f, ln), v in agg.items():
    byfile[f][0] += v[0]
    byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"{f:24s} inst {100 * v[0] / ti:5.1f}%  samples {100 * v[1] / max(ts, 1):5.1f}%")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
    print(f"{f}:{ln:<5d} inst {v[0]:9d} {100 * v[0] / ti:5.1f}%  samples {v[1]:6d} {100 * v[1] / max(ts, 1):5.1f}%")
