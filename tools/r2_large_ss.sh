#!/bin/bash
# round 2: slot-sliced vs frame-sliced kernel at c4 (2^16 frames) and c5 (2^15 frames)
cd "$(dirname "$0")/.."
for cfg in "c4 65536 4" "c5 32768 2"; do
  set -- $cfg
  for k in bs ss; do
    SCPD_KERNEL=$k python tools/quick_perf.py --cfg $1 --frames $2 --iters 3 --check $3 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-150 | sed "s/^/kernel=$k /"; echo
  done
done
