#!/bin/bash
# Round 2, GPU call 1 (1 GPU): default suite incl. the new loopback collective tests, the opt-in paths of round 1 under
# their switches, drivers with them on, bench default vs switches, reference arm.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call1.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
echo "=== default suite" >> $LOG
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 >> $LOG
echo "=== experimental: conv + bias + ELU (VAE / CPC), convT" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_CONV_ACT=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k "conv_bias_act or conv_transpose" 2>&1 | tail -12 >> $LOG
echo "=== experimental: fused BN backward" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_BN_BWD_FUSED=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k fused_bn 2>&1 | tail -8 >> $LOG
echo "=== experimental: fused residual-gradient accumulation" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_SKIP_FUSED=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k "identity_block or accumulating" 2>&1 | tail -12 >> $LOG
echo "=== experimental: fused classifier head" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_HEAD_FUSED=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k classifier_head 2>&1 | tail -8 >> $LOG
echo "=== drivers + kernels with ALL experimental paths on" >> $LOG
FEDB200_CONV_ACT=1 FEDB200_BN_BWD_FUSED=1 FEDB200_HEAD_FUSED=1 FEDB200_SKIP_FUSED=1 timeout 400 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 >> $LOG
echo "=== bench default" >> $LOG
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err >> $LOG; cut -c1-1500 gpurun_out/bench_default.json >> $LOG
for sw in "FEDB200_BN_BWD_FUSED=1" "FEDB200_HEAD_FUSED=1" "FEDB200_SKIP_FUSED=1" "FEDB200_BN_BWD_FUSED=1 FEDB200_HEAD_FUSED=1 FEDB200_SKIP_FUSED=1"; do
  echo "=== bench $sw" >> $LOG
  env $sw timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-collective-table 2>&1 | tail -1 | cut -c1-260 >> $LOG
done
echo "=== bench consensus --bb (config 3, N=1)" >> $LOG
timeout 300 python bench.py --driver consensus --bb --steps 20 --warmup 5 --no-e2e --no-collective-table 2>&1 | tail -1 | cut -c1-400 >> $LOG
echo "=== bench reference" >> $LOG
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-1200 >> $LOG
echo "=== done" >> $LOG
tail -60 $LOG
