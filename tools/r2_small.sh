#!/bin/bash
# round 2: where the slot-sliced kernel overtakes the int16x2 kernel as the batch grows (c1, c2, c3)
cd "$(dirname "$0")/.."
for cfg in c1 c2 c3; do
  for fr in 4096 8192 12288 16384 24576 32768; do
    [ $cfg = c3 ] && [ $fr -gt 16384 ] && continue
    for k in fast ss; do
      SCPD_KERNEL=$k python tools/quick_perf.py --cfg $cfg --frames $fr --iters 10 2>&1 | tail -1 | grep -o "info [0-9.]* Gb/s" | sed "s/^/$cfg frames=$fr kernel=$k /"
    done
  done
done
