#!/bin/bash
# round 2: ncu --set full of the slot-sliced walk at c1 (after a plain run of the same command)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/quick_perf.py --cfg c1 --frames 1048576 --iters 1"
$CMD > gpurun_out/plain_c1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sc_decode_ss" -s 2 -c 1 -f -o gpurun_out/prof_ss_c1_v7 $CMD > gpurun_out/ncu_c1.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/plain_c1.log | cut -c1-140
