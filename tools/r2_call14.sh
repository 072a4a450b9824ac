#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call14.log
: > $LOG
echo "=== dilated stem tests" >> $LOG
timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider --tb=short -k "dilated_stem" 2>&1 | tail -25 >> $LOG
echo "=== cpc driver test" >> $LOG
timeout 300 python -m pytest tests/test_gpu_drivers.py -m gpu -q -p no:cacheprovider --tb=short -k "cpc" 2>&1 | tail -8 >> $LOG
echo "=== diag_net2_tf32" >> $LOG
timeout 300 python tools/diag_net2_tf32.py 2>&1 | tail -30 >> $LOG
for d in cpc vae_cl vae; do
  echo "=== profile $d" >> $LOG
  timeout 400 python tools/profile_aux.py $d 2 2>&1 | tail -50 >> $LOG
done
echo "=== bench cpc" >> $LOG
timeout 600 python bench.py --driver cpc --steps 10 --warmup 3 2>&1 | tail -1 >> $LOG
echo "=== done" >> $LOG
