#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call3.log
: > $LOG
echo "=== wgrad tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_wgrad.py -m gpu -q -p no:cacheprovider --tb=line 2>&1 | tail -40 >> $LOG
echo "=== aux tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_aux.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -120 >> $LOG
echo "=== experimental (now default) tests" >> $LOG
timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -40 >> $LOG
echo "=== whole suite" >> $LOG
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -60 >> $LOG
echo "=== bench default x3" >> $LOG
for rep in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'time_to_first_step_s')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" >> $LOG 2>&1
done
echo "=== bench WGRAD=0" >> $LOG
FEDB200_WGRAD=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-collective-table 2>&1 | tail -1 | cut -c1-200 >> $LOG
echo "=== bench Net fedavg driver quick (default model) K=4" >> $LOG
timeout 300 python federated_multi.py --K 4 --Nloop 1 --Nadmm 1 --max_minibatches 20 --no-save_model --check_results 2>&1 | tail -8 >> $LOG
echo "=== done" >> $LOG
