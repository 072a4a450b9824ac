#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out/ncu
LOG=gpurun_out/r2_call23.log
: > $LOG
echo "=== launch list of one eager step" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2300 --launch-count 900 --csv --log-file gpurun_out/ncu/step_launches2.csv \
   python bench.py --gpus 1 --steps 3 --warmup 3 --no-graphs --no-e2e --no-collective-table > gpurun_out/ncu/step_launches2.out 2>&1
tail -1 gpurun_out/ncu/step_launches2.out | cut -c1-200 >> $LOG
wc -l gpurun_out/ncu/step_launches2.csv >> $LOG
echo "=== done" >> $LOG
