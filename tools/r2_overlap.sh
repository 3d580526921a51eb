#!/bin/bash
# round 2: Monte-Carlo loop with fewer walk warps per CTA, so that more channel CTAs fit beside the walk's (generator on its own stream)
cd "$(dirname "$0")/.."
for w in 16 14 12 10 8; do
  SCPD_SS_WARPS=$w python tools/r2_pipeline.py 2>&1 | tail -2 | sed "s/^/warps=$w /" | cut -c1-150
done
