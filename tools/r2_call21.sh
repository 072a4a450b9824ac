#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call21.log
: > $LOG
echo "=== bn tests" >> $LOG
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -5 | cut -c1-400 >> $LOG
echo "=== bench_bn" >> $LOG
timeout 300 python tools/bench_bn.py 2>&1 | grep -E "bwd" >> $LOG
echo "=== bench headline" >> $LOG
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 | cut -c1-1500 >> $LOG
echo "=== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -4 | cut -c1-400 >> $LOG
echo "=== done" >> $LOG
