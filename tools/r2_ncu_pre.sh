#!/bin/bash
# round 2: ncu --set full of the plane conversion (with the prefused f levels) and the slot-sliced walk at c2
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/quick_perf.py --cfg c2 --frames 1048576 --iters 1"
$CMD > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sc_decode_ss|ss_planes" -s 4 -c 2 -f -o gpurun_out/prof_ss_c2_pre $CMD > gpurun_out/ncu_c2.log 2>&1
echo "ncu rc=$?"
tail -2 gpurun_out/plain_c2.log
