#!/bin/bash
# A/B of run-time knobs of the slot-sliced kernel (development): "VAR=a,b,..." pairs, all configs given in CFGS
set -u
cd "$(dirname "$0")/.."
export SCPD_KERNEL=ss SCPD_SS_LWIN=${SCPD_SS_LWIN:-9}
O=gpurun_out/r2_ab_${TAG:-x}.txt
: > $O
CFGS=${CFGS:-c1 c2}
IFS=';' read -ra COMBOS <<< "$1"
for combo in "${COMBOS[@]}"; do
  for cfg in $CFGS; do
    case $cfg in c1|c2) F=1048576;; c3) F=131072;; c4) F=32768;; c5) F=16384;; esac
    echo "== $combo $cfg" >> $O
    env $combo timeout 300 python tools/quick_perf.py --cfg $cfg --frames $F --iters 3 --check 128 2>&1 | grep -E "Gb/s|check|rror" >> $O
  done
done
grep -E "==|Gb/s|MISMATCH|rror" $O | sed 's/prune=2 group=32//; s/coded.*//; s/N=.*frames\/s//' | paste - - | awk '{print $2,$3,$4,$5,$6,$NF, $(NF-1)}'
