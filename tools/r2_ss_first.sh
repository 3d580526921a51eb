#!/bin/bash
# round 2, first GPU contact of the slot-sliced kernel: parity spot checks + throughput + plan sweep
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2_first.txt
: > $O
run() { echo "== $*" >> $O; "$@" >> $O 2>&1; echo "rc=$?" >> $O; }
export SCPD_KERNEL=ss
run python tools/quick_perf.py --cfg c1 --frames 4096 --check 4096 --iters 2
run python tools/quick_perf.py --cfg c1 --frames 1048576 --check 2048
run python tools/quick_perf.py --cfg c2 --frames 1048576 --check 512
run python tools/quick_perf.py --cfg c3 --frames 131072 --check 64
run python tools/quick_perf.py --cfg c4 --frames 32768 --check 16
run python tools/quick_perf.py --cfg c5 --frames 16384 --check 4 --iters 2
for w in 8 12 16; do
  for lsa in 7 8 9; do
    SCPD_SS_WARPS=$w SCPD_SS_LSA=$lsa SCPD_VERBOSE=1 run python tools/quick_perf.py --cfg c1 --frames 1048576
    SCPD_SS_WARPS=$w SCPD_SS_LSA=$lsa run python tools/quick_perf.py --cfg c2 --frames 1048576
  done
done
SCPD_SS_FUSE=0 run python tools/quick_perf.py --cfg c1 --frames 1048576
unset SCPD_KERNEL
SCPD_KERNEL=bs run python tools/quick_perf.py --cfg c1 --frames 1048576
SCPD_KERNEL=bs run python tools/quick_perf.py --cfg c2 --frames 1048576
grep -E "Gb/s|check|rc=[1-9]|slot-sliced" $O
