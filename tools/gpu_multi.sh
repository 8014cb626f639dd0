#!/bin/bash
# Multi-GPU visit (under `gpurun --gpus N`): weak-scaled chain, strong-scaled configs[2] (64 frames over N ranks), training step.
N=${1:-8}; TAG=${2:-r02multi}
OUT=gpurun_out/$TAG; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout -s KILL 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_weak_${N}gpu.json 2> $OUT/bench_weak_${N}gpu.err; echo "weak rc=$?" | tee -a $OUT/status.txt
tail -c 600 $OUT/bench_weak_${N}gpu.json; tail -3 $OUT/bench_weak_${N}gpu.err
timeout -s KILL 600 $RUN bench.py --gpus $N --global-frames 64 --steps 10 --warmup 3 > $OUT/bench_strong64_${N}gpu.json 2> $OUT/bench_strong64_${N}gpu.err; echo "strong rc=$?" | tee -a $OUT/status.txt
tail -c 600 $OUT/bench_strong64_${N}gpu.json; tail -3 $OUT/bench_strong64_${N}gpu.err
timeout -s KILL 600 $RUN bench.py --gpus $N --workload train --steps 10 --warmup 3 > $OUT/bench_train_${N}gpu.json 2> $OUT/bench_train_${N}gpu.err; echo "train rc=$?" | tee -a $OUT/status.txt
tail -c 900 $OUT/bench_train_${N}gpu.json; tail -3 $OUT/bench_train_${N}gpu.err
cat $OUT/status.txt
