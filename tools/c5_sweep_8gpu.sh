#!/bin/bash
# BASELINE configs[4] with the final build on 8 GPUs of one box: the 131 072-frame sweep on 1 and 8 GPUs (strong scaling: 16 384
# frames per GPU and point at 8) and the sweep with 131 072 frames per GPU (weak scaling: 1 048 576 frames per point at 8)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/c5_sweep.sh "1 8" 131072 gpurun_out/c5_sweep_final_strong.jsonl > gpurun_out/c5_sweep_final.txt 2>&1
echo "--- weak: 131072 frames per GPU and point" >> gpurun_out/c5_sweep_final.txt
tools/c5_sweep.sh "8" 1048576 gpurun_out/c5_sweep_final_weak.jsonl >> gpurun_out/c5_sweep_final.txt 2>&1
cat gpurun_out/c5_sweep_final.txt
