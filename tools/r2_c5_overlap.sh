#!/bin/bash
# round 2: does the generator overlap the walk in the c5 Monte-Carlo loop when the walk leaves room? (warps per CTA of the walk)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
o=gpurun_out/r2_c5_overlap.txt
: > $o
python - <<'PY'
import sys
sys.path.insert(0, ".")
import sc_polar_decoder_hls_b200 as scpd
scpd.write_flags("/tmp/frozen_c5.txt", scpd.packed_flags("frozen_n_524288_k_262144", 524288))
PY
for w in 16 12 8; do
  echo "SCPD_SS_WARPS=$w" >> $o
  SCPD_SS_WARPS=$w SCPD_VERBOSE=1 timeout 300 tools/ber_bench --flags /tmp/frozen_c5.txt -n 524288 --snr 2:1:2 --frames 131072 --gpus 1 --json 2>&1 | cut -c1-400 | tail -4 >> $o
done
cat $o
