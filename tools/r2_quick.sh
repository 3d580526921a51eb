#!/bin/bash
# round 2: c1 / c2 / c3 device-resident throughput of the current build, each checked against the oracle
cd "$(dirname "$0")/.."
for cfg in "c1 1048576" "c2 1048576" "c3 131072"; do
  set -- $cfg
  python tools/quick_perf.py --cfg $1 --frames $2 --iters 5 --check 32 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-150; echo
done
