#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call16.log
: > $LOG
echo "=== new tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_aux.py tests/test_gpu_drivers.py -m gpu -q -p no:cacheprovider --tb=short -s -k "unfold or graphed_closure or vae_and_vae_cl_and_cpc or linear_f32" 2>&1 | grep -v Warning | tail -30 | cut -c1-1200 >> $LOG
for d in "--driver cpc" "--driver vae_cl" "--driver vae"; do
  echo "=== bench $d" >> $LOG
  timeout 900 python bench.py $d --steps 10 --warmup 3 --no-collective-table 2>&1 | tail -1 | cut -c1-400 >> $LOG
done
echo "=== profile cpc (graphs off)" >> $LOG
timeout 400 python tools/profile_aux.py cpc 2 2>&1 | grep -v Warn | head -24 | cut -c1-200 >> $LOG
echo "=== done" >> $LOG
