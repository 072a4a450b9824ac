#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call12.log
: > $LOG
for sel in "linear_f32" "act_bwd_bias" "maxpool2x2" "argmax_count" "info_nce" "gauss_nll" "small_direct_conv"; do
  echo "=== $sel + net" >> $LOG
  timeout 200 python -m pytest tests/test_gpu_aux.py -m gpu -q -p no:cacheprovider --tb=line -k "$sel or net_and_net2" 2>&1 | grep -E "passed|failed|conv1.weight" | cut -c1-300 >> $LOG
done
echo "=== smoke" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $LOG
echo "=== done" >> $LOG
