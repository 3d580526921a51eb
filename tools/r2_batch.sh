#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "run_ber" 2>&1 | tail -1
for mb in 512 1024 2048; do
  SCPD_BER_BATCH_MB=$mb python tools/r2_pipeline.py 2>&1 | tail -2 | sed "s/^/batch_mb=$mb /" | cut -c1-140
done
