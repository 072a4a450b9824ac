#!/bin/bash
# One gpurun call: per-test-function isolation (a sticky CUDA error must not mask the other results),
# then smoke(), then a short bench of both arms.  Logs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
OUT=gpurun_out/check.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $OUT 2>&1
python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))" >> $OUT 2>&1
TESTS=${TESTS:-"test_extension_is_native test_adam_prox test_adam_device test_vector_reductions test_lbfgs_two_loop test_lbfgs_on_cuda test_normalize_u8 test_linear_tf32 test_conv2d_nhwc test_bn_elu test_basic_block test_resnet18_fast test_cross_entropy test_fused_collective test_engine_resnet test_host_resident"}
for t in $TESTS; do
  echo "=== $t" >> $OUT
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$t" -p no:cacheprovider 2>&1 | tail -${TAIL:-25} >> $OUT
done
if [ -z "$SKIP_SMOKE" ]; then
  echo "=== smoke" >> $OUT
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 >> $OUT
fi
if [ -z "$SKIP_BENCH" ]; then
  echo "=== bench ours" >> $OUT
  timeout 900 python bench.py --gpus 1 --steps ${STEPS:-30} --warmup 5 2>&1 | tail -12 >> $OUT
  echo "=== bench ours (no fast, graphs)" >> $OUT
  timeout 900 python bench.py --gpus 1 --steps ${STEPS:-30} --warmup 5 --no-fast 2>&1 | tail -6 >> $OUT
  echo "=== bench reference" >> $OUT
  timeout 900 python bench.py --impl reference --gpus 1 --steps ${STEPS:-30} --warmup 5 2>&1 | tail -6 >> $OUT
fi
echo "=== done" >> $OUT
tail -5 $OUT
