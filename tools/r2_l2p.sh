#!/bin/bash
# round 2: the smallest workspace level in a dense array under an L2 persisting window (SCPD_SS_L2PERSIST), set-aside size
cd "$(dirname "$0")/.."
run() { python tools/quick_perf.py --cfg $1 --frames $2 --iters 5 2>&1 | grep "L2 persisting\|info" | cut -c1-130; }
for mb in 0 40 80; do
  echo "c2 set-aside >= $mb MB:"; SCPD_VERBOSE=1 SCPD_SS_L2PERSIST_MB=$mb run c2 1048576
done
echo "c2 window off:"; SCPD_SS_L2PERSIST=0 run c2 1048576
