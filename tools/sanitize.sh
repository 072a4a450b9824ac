#!/bin/bash
# compute-sanitizer targets for the hand-written kernels.  Run on a GPU box:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/sanitize.sh memcheck'
# tools: memcheck (out-of-bounds / misaligned accesses, incl. TMA-written shared memory), racecheck (shared-memory
# hazards in the epilogue staging tiles and reduction buffers), synccheck (barrier misuse), initcheck.
# The workload is the small-shape part of the kernel test-suite (every kernel family once); sanitizer slow-down is
# 20-100x, so the selection below stays under a few minutes.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL='test_adam_prox_matches_oracle or test_vector_reductions or test_bn_elu_forward_backward or test_cross_entropy_and_vae_loss or (test_conv2d_nhwc_tcgen05 and 4-32-64-64) or (test_conv_kernel_variants and 3-32-4-64) or (test_fused_collective_single_process_matches_torch and 5130) or test_stride2_data_gradient_as_one_stride1_conv'
timeout 850 compute-sanitizer --tool "$TOOL" --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:cacheprovider -k "$SEL" > "gpurun_out/sanitize_$TOOL.log" 2>&1
echo "exit code $?" >> "gpurun_out/sanitize_$TOOL.log"
tail -15 "gpurun_out/sanitize_$TOOL.log"
