#!/bin/bash
# round-2 development loop: throughput of the slot-sliced kernel at c1..c5 (+ optional ncu of c1 / c2 with a tag)
# usage: r2_perf.sh [tag] [configs...]     (SCPD_FAST_BUILD is honoured so that the box does not rebuild)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-}
shift || true
CFGS=${*:-c1 c2}
export SCPD_KERNEL=${SCPD_KERNEL:-ss}
O=gpurun_out/r2_perf_${TAG:-x}.txt
: > $O
for cfg in $CFGS; do
  case $cfg in c1|c2) F=1048576;; c3) F=131072;; c4) F=32768;; c5) F=16384;; esac
  python tools/quick_perf.py --cfg $cfg --frames $F --check 256 >> $O 2>&1
done
if [ -n "$TAG" ]; then
  for cfg in c1 c2; do
    CMD="python tools/quick_perf.py --cfg $cfg --frames 1048576 --iters 1"
    $CMD > gpurun_out/plain_$cfg.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:sc_decode_ss -s 2 -c 1 -f -o gpurun_out/prof_ss_${cfg}_$TAG $CMD > gpurun_out/ncu_$cfg.log 2>&1
    echo "ncu $cfg rc=$?" >> $O
  done
fi
grep -E "Gb/s|check|rc=" $O
