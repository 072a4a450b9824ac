#!/bin/bash
# First GPU call of the next round (single GPU, ~3 min): confirm the opt-in paths written after round 1's GPU budget
# was spent, then decide which defaults to flip.  Usage:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_validate.sh'
# Multi-GPU follow-up (2 GPUs): NGPU=2 bash tools/gpu_multi.sh ; torchrun --nproc-per-node 2 tools/bench_collective.py
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/next_round_validate.log
: > $LOG
echo "=== default suite" >> $LOG
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 >> $LOG
echo "=== experimental: conv + bias + ELU (VAE / CPC)" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_CONV_ACT=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k "conv_bias_act or conv_transpose" 2>&1 | tail -8 >> $LOG
echo "=== experimental: fused BN backward" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_BN_BWD_FUSED=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k fused_bn 2>&1 | tail -8 >> $LOG
echo "=== experimental: fused residual-gradient accumulation" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_SKIP_FUSED=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k "identity_block or accumulating" 2>&1 | tail -8 >> $LOG
echo "=== experimental: fused classifier head" >> $LOG
FEDB200_EXPERIMENTAL=1 FEDB200_HEAD_FUSED=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q -p no:cacheprovider -k classifier_head 2>&1 | tail -8 >> $LOG
echo "=== drivers with the experimental paths on" >> $LOG
FEDB200_CONV_ACT=1 FEDB200_BN_BWD_FUSED=1 FEDB200_HEAD_FUSED=1 timeout 300 python -m pytest tests/test_gpu_drivers.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 >> $LOG
echo "=== BN micro-benchmark, two-pass vs fused backward" >> $LOG
(cd tools && timeout 120 python bench_bn.py 2>&1 | grep bwd) >> $LOG
(cd tools && FEDB200_BN_BWD_FUSED=1 timeout 120 python bench_bn.py 2>&1 | grep bwd | sed 's/^/fused /') >> $LOG
echo "=== bench, default vs fused BN backward" >> $LOG
timeout 200 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-220 >> $LOG
FEDB200_BN_BWD_FUSED=1 timeout 200 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-220 >> $LOG
FEDB200_BN_BWD_FUSED=1 FEDB200_HEAD_FUSED=1 FEDB200_SKIP_FUSED=1 timeout 200 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-220 >> $LOG
tail -40 $LOG
