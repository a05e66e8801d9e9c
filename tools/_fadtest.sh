timeout 600 python -m pytest tests -m gpu -x -q -k "fad or frechet" 2>&1 | tail -3
for d in 512 768 1024; do timeout 120 python tools/fad_bench.py 510976 $d tcgen05 10 2>&1 | tail -1 | cut -c1-200; done
