#!/bin/bash
# variant x sync x fuse matrix at c1 / c2 (development)
set -u
cd "$(dirname "$0")/.."
export SCPD_KERNEL=ss SCPD_SS_LWIN=9
O=gpurun_out/r2_matrix.txt
: > $O
for lib in fg0 fg1; do
 for sync in 0 1 2 4; do
  for fuse in 1 2; do
   for cfg in c1 c2; do
     echo "== lib=$lib sync=$sync fuse=$fuse $cfg" >> $O
     SCPD_LIB_PATH=$PWD/sc_polar_decoder_hls_b200/variants/libscpd_$lib.so SCPD_SS_SYNC=$sync SCPD_SS_FUSE=$fuse timeout 120 python tools/quick_perf.py --cfg $cfg --frames 1048576 --iters 3 --check 256 2>&1 | grep -E "Gb/s|check|rror" >> $O
   done
  done
 done
done
cat $O
