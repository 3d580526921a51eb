#!/bin/bash
# round 2: ncu --set full of channel_kernel (c2, 2^18 frames, all-zero codeword), after a plain run of the same command
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/quick_perf.py --cfg c2 --frames 262144 --iters 1"
$CMD > gpurun_out/plain_chan.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:channel_kernel -c 1 -f -o gpurun_out/prof_chan_c2 $CMD > gpurun_out/ncu_chan.log 2>&1
echo "ncu rc=$?"
