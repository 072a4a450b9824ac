#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call22.log
: > $LOG
for v in "" "FEDB200_BN_BWD_FUSED=1" "FEDB200_PDL=1" "FEDB200_SKIP_FUSED=1"; do
  echo "=== bench headline $v" >> $LOG
  env $v timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --no-collective-table --no-e2e 2>&1 | tail -1 | cut -c1-260 >> $LOG
done
echo "=== launch list of one eager step" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 --launch-count 900 --csv --log-file gpurun_out/ncu/step_launches2.csv \
   python bench.py --gpus 1 --steps 3 --warmup 3 --no-graphs --no-e2e --no-collective-table > gpurun_out/ncu/step_launches2.out 2>&1
tail -1 gpurun_out/ncu/step_launches2.out | cut -c1-200 >> $LOG
wc -l gpurun_out/ncu/step_launches2.csv >> $LOG
echo "=== done" >> $LOG
