#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call15.log
: > $LOG
echo "=== new tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_aux.py tests/test_gpu_drivers.py -m gpu -q -p no:cacheprovider --tb=short -s -k "unfold or net2_fast or graphed_closure or vae_and_vae_cl_and_cpc or fedprox_lbfgs" 2>&1 | grep -v Warning | tail -40 | cut -c1-1500 >> $LOG
echo "=== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -15 | cut -c1-600 >> $LOG
for d in "--driver cpc" "--driver vae_cl" "--driver fedprox --optimizer lbfgs"; do
  echo "=== bench $d" >> $LOG
  timeout 900 python bench.py $d --steps 10 --warmup 3 --no-collective-table 2>&1 | tail -1 | cut -c1-1500 >> $LOG
done
echo "=== profile cpc (graphs off, after unfold)" >> $LOG
timeout 400 python tools/profile_aux.py cpc 2 2>&1 | grep -v Warn | head -14 >> $LOG
echo "=== done" >> $LOG
