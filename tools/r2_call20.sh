#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call20.log
: > $LOG
for r in 16 4 2; do
  echo "=== bench_bn FEDB200_RED_ROWS=$r" >> $LOG
  FEDB200_RED_ROWS=$r timeout 300 python tools/bench_bn.py 2>&1 | grep -E "bwd|col_stats" >> $LOG
done
echo "=== bn tests" >> $LOG
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short -k "bn_elu or resnet or col_stats" 2>&1 | tail -3 >> $LOG
for r in 16 4 2; do
  echo "=== bench headline FEDB200_RED_ROWS=$r" >> $LOG
  FEDB200_RED_ROWS=$r timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table --no-e2e 2>&1 | tail -1 | cut -c1-260 >> $LOG
done
echo "=== done" >> $LOG
