#!/bin/bash
# compute-sanitizer targets for the hand-written kernels.  Run on a GPU box:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh'            (all three tools)
# memcheck (out-of-bounds / misaligned accesses, incl. TMA-written shared memory; --report-api-errors no: the CUDA runtime's
# lazy module loading makes cuKernelGetFunction return CUDA_ERROR_INVALID_HANDLE once per kernel, which is not an error of ours),
# racecheck (shared-memory hazards in the epilogue staging tiles and reduction buffers; the opt-in CTA-pair variant
# FEDB200_2CTA is excluded: racecheck flags the smem word that tcgen05.alloc.cta_group::2 writes by hardware), synccheck.
# Workload: one small instance of every kernel family (sanitizer slow-down is 20-100x).
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
SEL='test_adam_prox_matches_oracle or test_vector_reductions or test_bn_elu_forward_backward or test_cross_entropy_and_vae_loss or (test_conv2d_nhwc_tcgen05 and 4-32-64-64) or (test_conv_kernel_variants and 3-32-4-64 and not 2CTA) or (test_fused_collective_single_process_matches_torch and 5130) or test_stride2_data_gradient_as_one_stride1_conv or (test_wgrad_matches_fp64_oracle and 5-6-20-12) or (test_wgrad_matches_fp64_oracle and 37-3-64-16) or test_act_bwd_bias or test_maxpool2x2 or test_argmax_count or (test_info_nce_fused and 16-8-3-3) or (test_small_direct_conv and 5-7-9-5) or (test_linear_f32 and 37-3-5) or (test_loopback_fedavg_fedprox_admm and 2-850) or (test_bb_update_kernel and 1)'
for TOOL in ${1:-memcheck racecheck synccheck}; do
  EXTRA=""
  [ "$TOOL" = "memcheck" ] && EXTRA="--report-api-errors no"
  timeout 700 compute-sanitizer --tool $TOOL $EXTRA --error-exitcode 7 --print-limit 20 \
     python -m pytest tests/test_gpu_kernels.py tests/test_gpu_wgrad.py tests/test_gpu_aux.py tests/test_gpu_loopback.py -m gpu -q -p no:cacheprovider -k "$SEL" > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitize_$TOOL.log
  echo "--- $TOOL"; grep -v "Host Frame" gpurun_out/sanitize_$TOOL.log | tail -6 | cut -c1-300
done
