#!/bin/bash
# round 2: slot-sliced (fused ops) vs frame-sliced kernel on the largest trees as the batch grows
cd "$(dirname "$0")/.."
for cfg in "c4 12288" "c4 16384" "c4 32768" "c5 12288" "c5 16384" "c5 24576"; do
  set -- $cfg
  for k in bs ss; do
    SCPD_KERNEL=$k python tools/quick_perf.py --cfg $1 --frames $2 --iters 3 2>&1 | tail -1 | grep -o "info [0-9.]* Gb/s" | sed "s/^/$1 frames=$2 kernel=$k /"
  done
done
