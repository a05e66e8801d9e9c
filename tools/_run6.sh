timeout 600 python -m pytest tests -m gpu -x -q -k "graphed or pipelined or steps_vs_reference" 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/bench_cfg2_host.json 2> gpurun_out/bench_cfg2_host.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg2_host.json"))
print("value", d["value"], d["ms_per_step"], "host", d["host_enqueue_ms_per_step"], "| e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "host", d["e2e"]["host_enqueue_ms_per_step"])
PY
