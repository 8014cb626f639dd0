#!/bin/bash
# One GPU visit: parity tests, smoke, bench, launch list, full ncu capture of our kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt
tail -c 3000 $OUT/bench.json
if [ "$2" != "quick" ]; then
  timeout 600 python tools/microbench.py > $OUT/microbench.jsonl 2> $OUT/microbench.err; echo "microbench rc=$?" | tee -a $OUT/status.txt
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu-launches rc=$?" | tee -a $OUT/status.txt
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k 'regex:roi_gather|confidence_reduce|lsq_|interp_fill|delaunay|extract_knots|loftr' -s 52 -c 40 -o $OUT/prof \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu-full rc=$?" | tee -a $OUT/status.txt
fi
tail -5 $OUT/pytest_gpu.log; tail -3 $OUT/smoke.log; cat $OUT/status.txt
