#!/bin/bash
# ncu --set full of the slot-sliced kernel at c1 and c2 (each after a plain run of the same command)
set -u
cd "$(dirname "$0")/.."
export SCPD_KERNEL=ss
TAG=${1:-v1}
for cfg in c1 c2; do
  CMD="python tools/quick_perf.py --cfg $cfg --frames 1048576 --iters 1"
  $CMD > gpurun_out/plain_$cfg.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sc_decode_ss -s 2 -c 1 -f -o gpurun_out/prof_ss_${cfg}_$TAG $CMD > gpurun_out/ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
  tail -2 gpurun_out/plain_$cfg.log
done
