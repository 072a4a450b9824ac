#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call19.log
: > $LOG
echo "=== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-800 >> $LOG
echo "=== bench headline (deferred rounds on / off)" >> $LOG
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table > gpurun_out/final_ours_n1.json 2>gpurun_out/final_ours_n1.err; tail -1 gpurun_out/final_ours_n1.json | cut -c1-1300 >> $LOG
FEDB200_DEFERRED_ROUNDS=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table --no-e2e 2>&1 | tail -1 | cut -c1-1300 >> $LOG
for d in "--driver vae" "--driver vae_cl" "--driver consensus --bb"; do
  echo "=== bench $d" >> $LOG
  timeout 900 python bench.py $d --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 | cut -c1-2500 >> $LOG
done
echo "=== done" >> $LOG
