#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call10.log
: > $LOG
echo "=== whole suite" >> $LOG
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -s 2>&1 | grep -v "^NCCL\|symm_mem" | tail -25 | cut -c1-400 >> $LOG
echo "=== Net2 diagnosis" >> $LOG
timeout 300 python tools/diag_net2.py 2>&1 | grep -v Warning | cut -c1-900 >> $LOG
echo "=== K=10 co-resident replicas: one stream vs one stream per replica" >> $LOG
timeout 600 python tools/k10_streams.py 2>&1 | grep streams >> $LOG
echo "=== bench fedprox + LBFGSNew (BASELINE config 4, N=1): ours, reference" >> $LOG
timeout 600 python bench.py --driver fedprox --optimizer lbfgs --steps 10 --warmup 3 --no-e2e --no-collective-table 2>&1 | tail -1 | cut -c1-700 >> $LOG
timeout 900 python bench.py --impl reference --driver fedprox --optimizer lbfgs --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-900 >> $LOG
echo "=== bench consensus --bb reference (config 3, N=1)" >> $LOG
timeout 900 python bench.py --impl reference --driver consensus --bb --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-700 >> $LOG
echo "=== final: reference then ours, N=1 (driver order)" >> $LOG
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref_n1.json 2>/dev/null; cut -c1-300 gpurun_out/final_ref_n1.json >> $LOG
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_ours_n1.json 2>/dev/null; cut -c1-300 gpurun_out/final_ours_n1.json >> $LOG
echo "=== done" >> $LOG
