#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call14.log
: > $LOG
echo "=== diag_net2_tf32" >> $LOG
timeout 300 python tools/diag_net2_tf32.py 2>&1 | tail -30 >> $LOG
for d in cpc vae_cl vae; do
  echo "=== profile $d" >> $LOG
  timeout 400 python tools/profile_aux.py $d 2 2>&1 | tail -50 >> $LOG
done
echo "=== done" >> $LOG
