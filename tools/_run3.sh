set -x
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 1500 gpurun_out/bench_cfg2.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
python bench.py --no-cpu-baseline --workload cfg3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 600 gpurun_out/bench_cfg3.json
python bench.py --no-cpu-baseline --workload cfg4 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 600 gpurun_out/bench_cfg4.json
python bench.py --no-cpu-baseline --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 600 gpurun_out/bench_cfg1.json
