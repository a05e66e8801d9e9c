set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gputest1.log; tail -5 gpurun_out/r2_gputest1.log
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_cfg2_a.json 2> gpurun_out/r2_bench_cfg2_a.err; tail -c 1500 gpurun_out/r2_bench_cfg2_a.json
python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_a.csv python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_warp -s 6 -c 1 -o gpurun_out/r2_prof_warp_a -f python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu_warp.log 2>&1
grep -i "stft_warp\|stft_pair" gpurun_out/r2_launches_a.csv | head -8
