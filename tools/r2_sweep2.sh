#!/bin/bash
# occupancy sweep: warps per CTA beyond 16 (library built with a larger SCPD_SS_THREADS)
set -u
cd "$(dirname "$0")/.."
export SCPD_KERNEL=ss
O=gpurun_out/r2_sweep2_${1:-x}.txt
: > $O
for cfg in c1 c2; do
for w in 16 20 24; do
  for lwin in 9 10; do
      echo "== $cfg warps=$w lsa=7 lwin=$lwin" >> $O
      SCPD_SS_WARPS=$w SCPD_SS_LSA=7 SCPD_SS_LWIN=$lwin SCPD_VERBOSE=1 python tools/quick_perf.py --cfg $cfg --frames 1048576 --iters 3 --check 256 2>&1 | grep -E "Gb/s|slot-sliced|check" >> $O
  done
done
done
cat $O
