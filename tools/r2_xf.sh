#!/bin/bash
# round 2: f / g fused with the child's opening f (SS_XF_*) in the slot-sliced kernel: c3 parity, then c2 ... c5
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "baseline_configs and c3 or every_kernel_variant and c3 and ss" 2>&1 | tail -2
run() { python tools/quick_perf.py --cfg $1 --frames $2 --iters 3 --check $3 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-150; echo; }
echo "c2 lean:";  SCPD_SS_XF_MIN_LOG2N=99 run c2 1048576 32
echo "c2 xf:";    SCPD_SS_XF_MIN_LOG2N=12 run c2 1048576 32
echo "c3 lean:";  SCPD_SS_XF_MIN_LOG2N=99 run c3 131072 8
echo "c3 xf:";    run c3 131072 8
echo "c4 bs:";    run c4 65536 4
echo "c4 ss xf:"; SCPD_KERNEL=ss run c4 65536 4
echo "c5 bs:";    run c5 32768 2
echo "c5 ss xf:"; SCPD_KERNEL=ss run c5 32768 2
