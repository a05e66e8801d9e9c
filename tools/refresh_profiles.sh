#!/bin/bash
# Reproduces the single-GPU evidence under profiles/r01 on a B200 box (run from the repo root, e.g. through gpurun):
#   bench lines of all four workloads + the reference arm, the ncu launch list of the eager step and one
#   `ncu --set full` capture of the dominant kernel (each profiler run only after the same command exited 0 plainly).
# Afterwards, here:  cp gpurun_out/{launches.csv,prof_pair.ncu-rep,bench_*.json} profiles/r01/ ;
#   ncu -i prof_pair.ncu-rep --page raw --csv > stft_pair_full_raw.csv ; tools/ncu_by_line.py ; tools/ncu_sass_summary.py
set -x
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 900 gpurun_out/bench_cfg2.json
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
export DM_STFT_FRAMES_PER_TILE=14
ncu --set full --clock-control none --import-source on -k regex:stft_pair -s 6 -c 1 -o gpurun_out/prof_pair -f python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu_pair.log 2>&1
unset DM_STFT_FRAMES_PER_TILE
for c in cfg1 cfg3 cfg4; do python bench.py --no-cpu-baseline --workload $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; done
python - <<PY
import json
for c in ("cfg1","cfg2","cfg3","cfg4"):
    d=json.load(open(f"gpurun_out/bench_{c}.json"))
    print(c, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4), "dom", d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
PY
