#!/bin/bash
# plan sweep of the slot-sliced kernel: warps per CTA x resident alpha levels x partial-sum window
set -u
cd "$(dirname "$0")/.."
export SCPD_KERNEL=ss
O=gpurun_out/r2_sweep.txt
: > $O
for cfg in c2 c1; do
for w in 8 10 12 14 16; do
  for lsa in 7 8; do
    for lwin in 8 10; do
      echo "== $cfg warps=$w lsa=$lsa lwin=$lwin" >> $O
      SCPD_SS_WARPS=$w SCPD_SS_LSA=$lsa SCPD_SS_LWIN=$lwin SCPD_VERBOSE=1 python tools/quick_perf.py --cfg $cfg --frames 1048576 --iters 3 2>&1 | grep -E "Gb/s|slot-sliced" >> $O
    done
  done
done
done
cat $O
