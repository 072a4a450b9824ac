#!/bin/bash
# shuffle-store dgrad: kernel tests, full GPU suite, A/B bench
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call13.log
: > $LOG
echo "=== stride2 kernel tests" >> $LOG
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short -k "stride2" 2>&1 | tail -15 >> $LOG
echo "=== net tests (verbose output)" >> $LOG
timeout 300 python -m pytest tests/test_gpu_aux.py -m gpu -q -p no:cacheprovider --tb=line -s -k "net_fast_path or net2_fast" 2>&1 | grep -E "passed|failed|Net2 grad|conv1.weight" | cut -c1-900 >> $LOG
echo "=== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -25 >> $LOG
echo "=== bench shuffle on" >> $LOG
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 >> $LOG
echo "=== bench shuffle off" >> $LOG
FEDB200_SHUFFLE_STORE=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 >> $LOG
echo "=== done" >> $LOG
