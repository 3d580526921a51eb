timeout 200 python -m pytest tests -m gpu -x -q -k "every_kernel_variant or ragged or golden or arbitrary or baseline_configs" 2>&1 | tail -3
q() { env "$@" python tools/quick_perf.py --cfg $CFG --frames $FR --iters 3 --check 16 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-150; echo; }
for spec in "c5 1024" "c5 256" "c4 1024" "c3 1024" "c2 512"; do
  set -- $spec; export CFG=$1 FR=$2
  echo "[auto]"; q SCPD_X=0
  echo "[fast32]"; q SCPD_KERNEL=fast SCPD_GROUP=32
  echo "[coop4]"; q SCPD_KERNEL=fast SCPD_COOP=4
  echo "[coop8]"; q SCPD_KERNEL=fast SCPD_COOP=8
done
