#!/bin/bash
# Round 2, GPU call 2 (1 GPU): wgrad kernel tests, diagnosis of the all-switches failure, benches after the warm-up fix.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call2.log
: > $LOG
echo "=== wgrad tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_wgrad.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -60 >> $LOG
echo "=== all switches, stop at first failure" >> $LOG
FEDB200_CONV_ACT=1 FEDB200_BN_BWD_FUSED=1 FEDB200_HEAD_FUSED=1 FEDB200_SKIP_FUSED=1 timeout 400 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider --tb=long 2>&1 | tail -80 >> $LOG
for sw in "FEDB200_CONV_ACT=1" "FEDB200_SKIP_FUSED=1" "FEDB200_BN_BWD_FUSED=1" "FEDB200_HEAD_FUSED=1"; do
  echo "=== drivers with $sw" >> $LOG
  env $sw timeout 400 python -m pytest tests/test_gpu_drivers.py -m gpu -q -p no:cacheprovider --tb=line 2>&1 | tail -8 >> $LOG
done
echo "=== default drivers + kernels (wgrad on)" >> $LOG
timeout 400 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 >> $LOG
for sw in "FEDB200_X=0" "FEDB200_WGRAD=0" "FEDB200_BN_BWD_FUSED=1" "FEDB200_HEAD_FUSED=1" "FEDB200_SKIP_FUSED=1" "FEDB200_BN_BWD_FUSED=1 FEDB200_HEAD_FUSED=1 FEDB200_SKIP_FUSED=1"; do
  for rep in 1 2; do
    echo "=== bench $sw ($rep)" >> $LOG
    env $sw timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-collective-table 2>&1 | tail -1 | cut -c1-200 >> $LOG
  done
done
echo "=== done" >> $LOG
