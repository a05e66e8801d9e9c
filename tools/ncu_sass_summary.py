"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): instructions executed and stall samples per opcode.

    ncu -i rep.ncu-rep --page source --csv > sass.csv ; python tools/ncu_sass_summary.py sass.csv
"""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
by_op = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r in rows[hdr_i + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break  # first launch only
    if len(r) <= max(ci, cs):
        continue
    src = r[ct].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "BAR")) and "." in op else "")
    n, s = int(r[ci] or 0), int(r[cs] or 0)
    by_op[op][0] += n
    by_op[op][1] += s
    by_op[op][2] += 1
    tot_i += n
    tot_s += s
print(f"total warp instructions {tot_i}, samples {tot_s}")
for op, (n, s, k) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{op:14s} inst {n:10d} {100 * n / tot_i:5.1f}%   samples {s:7d} {100 * s / max(tot_s, 1):5.1f}%   static {k}")
