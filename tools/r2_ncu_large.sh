#!/bin/bash
# round 2: ncu --set full at c3 (slot-sliced kernel, 2^17 frames) and c5 (frame-sliced kernel, 2^14 frames), each after a plain run
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
C=${1:-c3}; F=${2:-131072}
CMD="python tools/quick_perf.py --cfg $C --frames $F --iters 1"
$CMD > gpurun_out/plain_$C.log 2>&1 &&
ncu --set full --clock-control none -k regex:"sc_decode_|_planes" -s 4 -c 2 -f -o gpurun_out/prof_$C $CMD > gpurun_out/ncu_$C.log 2>&1
echo "$C ncu rc=$?"; tail -1 gpurun_out/plain_$C.log | cut -c1-150
