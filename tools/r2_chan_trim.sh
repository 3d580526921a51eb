#!/bin/bash
# round 2: channel fast path without the denormal pre-scaling of __logf / rsqrtf, guard as a kernel argument
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
o=gpurun_out/r2_chan_trim.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "channel or run_ber" 2>&1 | tail -4 > $o
timeout 300 python tools/r2_pipeline.py c1 c2 >> $o 2>&1
timeout 300 tools/probe/chan_err >> $o 2>&1
cat $o
