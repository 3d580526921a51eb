#!/bin/bash
# Small-batch throughput by kernel family and lane-group width (development aid; results in profiles/tuning_r1.md)
run() { # cfg frames label env...
  local cfg=$1 frames=$2 label=$3; shift 3
  env "$@" python tools/quick_perf.py --cfg $cfg --frames $frames --iters 3 --check 16 2>&1 | tail -2 | tr '\n' ' ' | sed "s/^/[$label] /" | cut -c1-175; echo
}
for spec in "c5 1024" "c5 4096" "c4 4096" "c4 8192" "c2 2048" "c2 24576" "c1 2048"; do
  set -- $spec
  run $1 $2 auto SCPD_X=0
  [ -n "$WITH_G8" ] && run $1 $2 fast8 SCPD_KERNEL=fast SCPD_GROUP=8
  run $1 $2 fast16 SCPD_KERNEL=fast SCPD_GROUP=16
  run $1 $2 fast32 SCPD_KERNEL=fast SCPD_GROUP=32
  run $1 $2 bs SCPD_KERNEL=bs
done
