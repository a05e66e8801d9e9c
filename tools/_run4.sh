timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/bench_cfg2_rs.json 2> gpurun_out/bench_cfg2_rs.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg2_rs.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"])
PY
python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
grep -c resample2 gpurun_out/launches.csv
