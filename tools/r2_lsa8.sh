#!/bin/bash
# round 2: 14 warps per CTA with alpha levels 6-8 in shared memory (level 8 no longer in the workspace), level 9 in tensor memory
cd "$(dirname "$0")/.."
run() { python tools/quick_perf.py --cfg $1 --frames 1048576 --iters 5 --check 32 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-140; echo; }
for c in c1 c2; do
  echo "$c default:"; run $c
  echo "$c w14 lsa8 lwin9:"; SCPD_SS_WARPS=14 SCPD_SS_LSA=8 SCPD_SS_LWIN=9 run $c
  echo "$c w14 lsa8 lwin8:"; SCPD_SS_WARPS=14 SCPD_SS_LSA=8 SCPD_SS_LWIN=8 run $c
  echo "$c w13 lsa8 lwin10:"; SCPD_SS_WARPS=13 SCPD_SS_LSA=8 SCPD_SS_LWIN=10 run $c
done
