"""Summarise an ncu CSV of the bandwidth-bound kernels (one row per kernel name: launches, median duration, DRAM bytes,
achieved DRAM GB/s and its share of the measured copy peak).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'ew_update|norm_update|mask_apply|add_scaled|fold_adjoint|resample|residual_wav|rir_' --csv \
        --log-file gpurun_out/bw.csv python tools/operator_bench.py --batch 128 --iters 3
    python tools/ncu_bandwidth_summary.py gpurun_out/bw.csv > profiles/r02/bandwidth_kernels_b128.md
"""
import collections
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6541.8
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ck, cm, cu, cv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
cid = hdr.index("ID")
per = collections.defaultdict(dict)
name = {}
for r in rows[hi + 1:]:
    if len(r) <= cv:
        continue
    v = float(r[cv].replace(",", ""))
    u = r[cu]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[r[cid]][r[cm]] = v * scale
    name[r[cid]] = r[ck]
agg = collections.defaultdict(list)
for i, m in per.items():
    if "gpu__time_duration.sum" in m:
        agg[name[i]].append((m["gpu__time_duration.sum"], m.get("dram__bytes_read.sum", 0.0) +
                             m.get("dram__bytes_write.sum", 0.0)))
print(f"| kernel | launches | median us | DRAM MB / launch | DRAM GB/s | % of {peak:.0f} GB/s |")
print("|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -statistics.median(t for t, _ in kv[1])):
    t = statistics.median(x for x, _ in v)
    b = statistics.median(x for _, x in v)
    gbs = b / (t * 1e-6) / 1e9 if t > 0 else 0.0
    short = k.split("(")[0].replace("void ", "")
    print(f"| `{short}` | {len(v)} | {t:.2f} | {b / 1e6:.2f} | {gbs:.0f} | {100 * gbs / peak:.1f} |")
