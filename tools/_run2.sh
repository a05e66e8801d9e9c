set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_gputest4.log; tail -8 gpurun_out/r2_gputest4.log
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_cfg2_d.json 2> gpurun_out/r2_bench_cfg2_d.err; tail -c 600 gpurun_out/r2_bench_cfg2_d.json
python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_d.csv python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_warp -s 6 -c 1 -o gpurun_out/r2_prof_warp_d -f python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu_warp.log 2>&1
grep -i "stft_warp" gpurun_out/r2_launches_d.csv | head -4
for c in cfg3 cfg4; do python bench.py --no-cpu-baseline --workload $c > gpurun_out/r2_bench_${c}_d.json 2> gpurun_out/r2_bench_${c}_d.err; tail -c 300 gpurun_out/r2_bench_${c}_d.json; done
