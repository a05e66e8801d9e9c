python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python tools/operator_bench.py --batch 16 > gpurun_out/operator_bench_b16.json 2> gpurun_out/operator_bench.err || tail -20 gpurun_out/operator_bench.err
python tools/operator_bench.py --batch 128 > gpurun_out/operator_bench_b128.json 2>> gpurun_out/operator_bench.err || tail -20 gpurun_out/operator_bench.err
python - <<PY
import json
for b in (16,128):
    d=json.load(open(f"gpurun_out/operator_bench_b{b}.json"))
    print("batch",b)
    for r in d["rows"]: print(f"  {r['op']:34s} {r['ms']*1000:8.1f} us {r['algorithmic_MB']:8.1f} MB {r['GBps']:8.0f} GB/s {100*r['frac_of_hbm_peak']:5.1f}%")
PY
