#!/bin/bash
# round 2: prefused f levels at N = 2048 and at N = 32768 (slot-sliced kernel forced) against the defaults
cd "$(dirname "$0")/.."
for pre in 0 1 2; do
  SCPD_SS_PRE=$pre python tools/quick_perf.py --cfg n2048 --frames 1048576 --check 32 2>&1 | tail -1 | cut -c1-120 | sed "s/^/pre=$pre /"
done
python tools/quick_perf.py --cfg c3 --frames 131072 --check 8 2>&1 | tail -1 | cut -c1-120 | sed "s/^/default /"
for pre in 0 2 3; do
  SCPD_KERNEL=ss SCPD_SS_PRE=$pre python tools/quick_perf.py --cfg c3 --frames 131072 --check 8 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-150 | sed "s/^/ss pre=$pre /"; echo
done
