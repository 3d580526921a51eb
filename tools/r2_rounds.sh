#!/bin/bash
# round 2: batches in whole rounds of the resident warps (148 SMs x 16 warps x 32 frames = 75 776 frames) at c3 / c4 / c5
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
o=gpurun_out/r2_rounds.txt
: > $o
run() { timeout 300 python tools/quick_perf.py "$@" 2>&1 | grep -v "^$" | tail -2 >> $o; nvidia-smi --query-gpu=memory.used --format=csv,noheader >> $o; }
run --cfg c3 --frames 131072 --iters 5
run --cfg c3 --frames 151552 --iters 5
run --cfg c3 --frames 227328 --iters 5
run --cfg c4 --frames 65536 --iters 3
run --cfg c4 --frames 75776 --iters 3
run --cfg c4 --frames 151552 --iters 3
run --cfg c5 --frames 32768 --iters 2
run --cfg c5 --frames 65536 --iters 2 --check 2
run --cfg c5 --frames 75776 --iters 2
cat $o
