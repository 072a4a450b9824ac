#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call11.log
: > $LOG
timeout 300 python tools/diag_net_seeds.py 2>&1 | grep -v Warn | cut -c1-400 >> $LOG
echo "=== lbfgs tests after the batched-read change" >> $LOG
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_drivers.py -m gpu -q -p no:cacheprovider --tb=short -k "lbfgs or vae or cpc" 2>&1 | tail -5 >> $LOG
echo "=== aux benches again (vae_cl, cpc)" >> $LOG
for d in vae_cl cpc; do timeout 300 python bench.py --driver $d --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-200 >> $LOG; done
echo "=== sanitizers (clean configuration)" >> $LOG
bash tools/sanitize.sh >> $LOG 2>&1
echo "=== done" >> $LOG
