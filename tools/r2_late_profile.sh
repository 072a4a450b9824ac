#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out/ncu
LOG=gpurun_out/r2_late_profile.log
: > $LOG
echo "=== ncu --set full: late-round kernel modes" >> $LOG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"igemm_persistent|gemm_f32|convT_pack|wgrad_tf32" \
   --launch-count 30 -o gpurun_out/ncu/late_full -f python tools/ncu_late.py > gpurun_out/ncu/late_full.out 2>&1
tail -2 gpurun_out/ncu/late_full.out | cut -c1-200 >> $LOG
if [ -f gpurun_out/ncu/late_full.ncu-rep ]; then
  ncu -i gpurun_out/ncu/late_full.ncu-rep --page raw --csv > gpurun_out/ncu/late_full_raw.csv 2>/dev/null
  ls -la gpurun_out/ncu/late_full.ncu-rep gpurun_out/ncu/late_full_raw.csv >> $LOG
  if [ $(stat -c %s gpurun_out/ncu/late_full.ncu-rep) -gt 20000000 ]; then rm -f gpurun_out/ncu/late_full.ncu-rep; fi
fi
echo "=== compute-sanitizer over the same launches" >> $LOG
for TOOL in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $TOOL --report-api-errors no --error-exitcode 7 --print-limit 20 python tools/ncu_late.py > gpurun_out/sanitize_late_$TOOL.log 2>&1
  echo "$TOOL exit code $?" >> $LOG
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK" gpurun_out/sanitize_late_$TOOL.log | tail -2 >> $LOG
done
echo "=== consensus test + bench (ADMM deferred rounds)" >> $LOG
timeout 600 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_loopback.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -4 | cut -c1-400 >> $LOG
timeout 600 python bench.py --driver consensus --bb --steps 20 --warmup 5 --no-collective-table --no-e2e 2>&1 | tail -1 | cut -c1-1300 >> $LOG
echo "=== done" >> $LOG
