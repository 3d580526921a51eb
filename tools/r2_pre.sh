#!/bin/bash
# round 2: leading f levels computed by the plane conversion (SCPD_SS_PRE = 0 .. 3): parity, then c1 / c2 throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slot_sliced or (every_kernel_variant and ss)" 2>&1 | tail -3
for pre in 0 1 2 3; do
  for c in c1 c2; do
    SCPD_SS_PRE=$pre python tools/quick_perf.py --cfg $c --frames 1048576 --check 64 --iters 5 2>&1 | tail -2 | sed "s/^/pre=$pre /"
  done
done
