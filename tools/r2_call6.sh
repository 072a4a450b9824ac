#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call6.log
: > $LOG
echo "=== Net per-parameter errors" >> $LOG
timeout 300 python tools/diag_net.py >> $LOG 2>&1
echo "=== wgrad tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_wgrad.py -m gpu -q -p no:cacheprovider --tb=line 2>&1 | tail -8 >> $LOG
echo "=== wgrad micro-benchmark (TMA reduce epilogue, one-wave splits)" >> $LOG
timeout 600 python tools/bench_wgrad.py >> $LOG 2>&1
echo "=== same, element reds" >> $LOG
FEDB200_WGRAD_TMA_RED=0 timeout 600 python tools/bench_wgrad.py >> $LOG 2>&1
echo "=== bench" >> $LOG
timeout 300 python bench.py --steps 20 --warmup 5 --no-collective-table --no-e2e 2>&1 | tail -1 | cut -c1-200 >> $LOG
echo "=== done" >> $LOG
