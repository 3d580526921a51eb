#!/bin/bash
# round 2: the Monte-Carlo loop with all ten counters in one pass over x^ (count_all_kernel / count_all_smem_kernel);
# tests of the new kernels first, then the loop at c1 / c2 and the c5 loop of BASELINE configs[4] on one GPU
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
o=gpurun_out/r2_count.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "run_ber or extract_info or channel" 2>&1 | tail -25 > $o
timeout 300 python tools/r2_pipeline.py c1 c2 c3 c4 >> $o 2>&1
timeout 600 tools/c5_sweep.sh "1" 131072 gpurun_out/c5_sweep_1gpu.jsonl >> $o 2>&1
cat $o
