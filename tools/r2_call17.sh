#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call17.log
: > $LOG
echo "=== conv kernel tests (tap packing on)" >> $LOG
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_experimental.py tests/test_gpu_wgrad.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | grep -v Warning | tail -25 | cut -c1-1200 >> $LOG
echo "=== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-600 >> $LOG
for d in "--driver vae" "--driver cpc" "--driver vae_cl"; do
  for tp in 1 0; do
    echo "=== bench $d TAP_PACK=$tp" >> $LOG
    FEDB200_TAP_PACK=$tp timeout 900 python bench.py $d --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 | cut -c1-330 >> $LOG
  done
done
echo "=== bench headline" >> $LOG
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 | cut -c1-330 >> $LOG
FEDB200_TAP_PACK=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table --no-e2e 2>&1 | tail -1 | cut -c1-330 >> $LOG
echo "=== profile vae" >> $LOG
timeout 400 python tools/profile_aux.py vae 2 2>&1 | grep -v Warn | head -16 | cut -c1-200 >> $LOG
echo "=== done" >> $LOG
