timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stft_kernels_agree or bit_reproducible or loss_and_vjp" 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_pair.log
cat gpurun_out/pytest_pair.log
for nf in 8 10 12 14 15; do
  DM_STFT_FRAMES_PER_TILE=$nf timeout 300 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_pair_nf$nf.json 2>gpurun_out/bench_pair_nf$nf.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pair_nf$nf.json"))
print($nf, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"])
PY
done
