#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call18.log
: > $LOG
echo "=== conv kernel tests (tap packing on)" >> $LOG
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=line -k "conv" 2>&1 | grep -v Warning | tail -12 | cut -c1-600 >> $LOG
timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider --tb=line 2>&1 | grep -v Warning | tail -12 | cut -c1-600 >> $LOG
echo "=== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-600 >> $LOG
for d in "--driver vae" "--driver vae_cl"; do
  for tp in 1 0; do
    echo "=== bench $d TAP_PACK=$tp" >> $LOG
    FEDB200_TAP_PACK=$tp timeout 900 python bench.py $d --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 | cut -c1-330 >> $LOG
  done
done
echo "=== profile vae" >> $LOG
timeout 400 python tools/profile_aux.py vae 2 2>&1 | grep -v Warn | head -18 | cut -c1-200 >> $LOG
echo "=== profile vae_cl" >> $LOG
timeout 400 python tools/profile_aux.py vae_cl 2 2>&1 | grep -v Warn | head -18 | cut -c1-200 >> $LOG
echo "=== done" >> $LOG
