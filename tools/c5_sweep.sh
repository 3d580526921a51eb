#!/bin/bash
# BASELINE configs[4]: N = 524288, K = 262144 sharded over 1 / 2 / 4 / 8 GPUs of one box, BER / FER sweep 1 .. 4 dB.
# Frames are split over the GPUs as contiguous ranges of one stream (each GPU jumps the generators to its first frame);
# the only exchange is the sum of the counters on the host.  One JSON line per (GPU count, Eb/N0) point.
# usage: c5_sweep.sh "1 2 4 8" [frames per point] [out file]
set -eu
cd "$(dirname "$0")/.."
GPUS=${1:-1}
FRAMES=${2:-131072}
OUT=${3:-gpurun_out/c5_sweep.jsonl}
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import sc_polar_decoder_hls_b200 as scpd
n = 524288
scpd.write_flags("/tmp/frozen_c5.txt", scpd.packed_flags("frozen_n_524288_k_262144", n))
PY
: > "$OUT"
for g in $GPUS; do
  tools/ber_bench --flags /tmp/frozen_c5.txt -n 524288 --snr 1:0.5:4 --frames "$FRAMES" --gpus "$g" --json >> "$OUT"
done
python - "$OUT" <<'PY'
import json, sys, collections
rows = [json.loads(l) for l in open(sys.argv[1])]
by = collections.defaultdict(dict)
for r in rows:
    by[r["ebn0_db"]][r["gpus"]] = r
ok = True
for snr, d in sorted(by.items()):
    ref = d[min(d)]
    line = f"Eb/N0 {snr:4.2f} dB  BER {ref['bit_errors']/ref['bits']:.3e} FER {ref['frame_errors']/ref['frames']:.3e} |"
    for g, r in sorted(d.items()):
        same = (r["bit_errors"], r["frame_errors"], r["frames"]) == (ref["bit_errors"], ref["frame_errors"], ref["frames"])
        ok &= same
        line += f" {g} GPU: {r['info_gbps_incl_channel']:7.1f} Gb/s{'' if same else ' COUNTERS DIFFER'}"
    print(line)
print("summed counters equal the 1-GPU counters at every point" if ok else "MISMATCH")
PY
