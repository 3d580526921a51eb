#!/bin/bash
# round 2: frame-sliced kernel at c5 with 2^14 frames (512 groups for 148 SMs): warps per CTA x shared memory per group
cd "$(dirname "$0")/.."
for w in 4 2 1; do
  for kb in 7 15 31 63; do
    SCPD_BS_WARPS=$w SCPD_BS_SMEM_KB=$kb python tools/quick_perf.py --cfg c5 --frames 16384 --iters 3 --check 2 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-160 | sed "s/^/warps=$w smem_kb=$kb /"; echo
  done
done
