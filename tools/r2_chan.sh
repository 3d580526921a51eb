#!/bin/bash
# round 2: guarded fast path of the channel kernel: same LLRs (test), kernel time per mode, Monte-Carlo pipeline
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "channel or run_ber" 2>&1 | tail -3
for m in 0 2 1; do
  SCPD_CHANNEL_FAST=$m python tools/r2_pipeline.py 2>&1 | tail -4 | sed "s/^/mode=$m /"
done
