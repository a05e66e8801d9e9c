#!/bin/bash
# Reproduces the single-GPU evidence under profiles/r02 on a B200 box (run from the repo root, e.g. through gpurun):
#   bench lines of all four step workloads + the reference arm, the ncu launch list of the eager step, one
#   `ncu --set full` capture of the dominant kernel, the bandwidth table of the streaming kernels at 128 clips and the
#   A/B sweeps of the dm_set_tuning knobs (each profiler run only after the same command exited 0 plainly).
# Afterwards, here:
#   cp gpurun_out/{bench_*.json,launches_cfg2_eager.csv,kernel_sweep_*.json,operator_bench_b128.json} profiles/r02/
#   python tools/ncu_bandwidth_summary.py gpurun_out/bw_b128.csv > profiles/r02/bandwidth_kernels_b128.md
#   ncu -i gpurun_out/prof_warp.ncu-rep --page raw --csv > profiles/r02/stft_warp_full_raw.csv
#   ncu -i gpurun_out/prof_warp.ncu-rep --page source --csv > /tmp/sass.csv ; cuobjdump -xelf all <lib>.so ;
#   nvdisasm -g -c stft_warp.sm_100a.cubin > /tmp/all.sass ; python tools/ncu_by_line.py /tmp/all.sass <mangled> /tmp/sass.csv 40 samples
# Multi-GPU lines: `gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr
#   127.0.0.1 --master-port 29541 bench.py --gpus N [--workload cfg5 --fad-check]`.
set -x
python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
python bench.py > gpurun_out/bench_cfg2_graph.json 2> gpurun_out/bench_cfg2.err; tail -c 900 gpurun_out/bench_cfg2_graph.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
for c in cfg1 cfg3 cfg4; do python bench.py --no-cpu-baseline --workload $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; done
python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg2_eager.csv python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_warp -s 6 -c 1 -o gpurun_out/prof_warp -f python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu_warp.log 2>&1
python tools/operator_bench.py --batch 128 --iters 10 > gpurun_out/operator_bench_b128.json 2> gpurun_out/opb.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ew_update|norm_update|mask_apply|add_scaled|fold_adjoint|resample|residual_wav|rir_' --csv --log-file gpurun_out/bw_b128.csv python tools/operator_bench.py --batch 128 --iters 3 > gpurun_out/ncu_bw.log 2>&1
for B in 16 128; do python tools/kernel_sweep.py --batch $B --iters 12 > gpurun_out/kernel_sweep_b$B.json 2> gpurun_out/sweep.err; done
for P in 0 1; do DM_TUNE_PDL=$P python bench.py --no-cpu-baseline > gpurun_out/bench_cfg2_pdl$P.json 2>/dev/null; done
python - <<PY
import json
for c in ("cfg2_graph","cfg1","cfg3","cfg4"):
    d=json.load(open(f"gpurun_out/bench_{c}.json"))
    print(c, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4), "dom", d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
PY
