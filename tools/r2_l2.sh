#!/bin/bash
# round 2: the reference-pruning mode (SCPD_PRUNE_REF_LEVEL2) on the GPU: parity tests, then throughput at c1 / c2 / c3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reference_pruning_level2 or raw_kernel" 2>&1 | tail -5
for c in c1 c2; do
  python tools/quick_perf.py --cfg $c --frames 262144 --prune 3 --check 64 2>&1 | tail -2
  SCPD_KERNEL=raw python tools/quick_perf.py --cfg $c --frames 262144 --prune 1 --check 64 2>&1 | tail -2
done
python tools/quick_perf.py --cfg c3 --frames 8192 --prune 3 --check 8 2>&1 | tail -2
