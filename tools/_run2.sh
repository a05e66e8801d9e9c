export DM_STFT_FRAMES_PER_TILE=${NF:-14}
python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:stft_pair -s 6 -c 1 -o gpurun_out/prof_pair -f python bench.py --no-cpu-baseline --eager --steps 3 --warmup 3 > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log
