#!/bin/bash
# tensor-memory level sweep
set -u
cd "$(dirname "$0")/.."
export SCPD_KERNEL=ss
O=gpurun_out/r2_sweep3.txt
: > $O
for cfg in c1 c2; do
for w in 12 16; do
 for lsa in 7 8; do
  for ltm in 0 8 9; do
      [ $ltm -le $lsa ] && [ $ltm -ne 0 ] && continue
      echo "== $cfg warps=$w lsa=$lsa lwin=9 ltm=$ltm" >> $O
      SCPD_SS_WARPS=$w SCPD_SS_LSA=$lsa SCPD_SS_LWIN=9 SCPD_SS_LTM=$ltm SCPD_VERBOSE=1 timeout 120 python tools/quick_perf.py --cfg $cfg --frames 1048576 --iters 3 --check 256 2>&1 | grep -E "Gb/s|slot-sliced|check|rror" >> $O
  done
 done
done
done
cat $O
