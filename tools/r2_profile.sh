#!/bin/bash
# Round 2 evidence run (1 GPU): Net diagnosis, ncu launch list of one eager step, ncu --set full captures (converted to CSV on
# the box; the .ncu-rep of the two headline kernels is kept), compute-sanitizer logs, accuracy / rounds-to-target run.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out/ncu
LOG=gpurun_out/r2_profile.log
: > $LOG
echo "=== Net test in isolation" >> $LOG
timeout 300 python -m pytest tests/test_gpu_aux.py -m gpu -q -p no:cacheprovider --tb=short -k "net_and_net2" 2>&1 | tail -12 | cut -c1-1500 >> $LOG
echo "=== launch list of one eager training step (device time per kernel, serialised, cold cache: compare shares)" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 --launch-count 900 --csv --log-file gpurun_out/ncu/step_launches.csv \
   python bench.py --gpus 1 --steps 3 --warmup 3 --no-graphs --no-e2e --no-collective-table > gpurun_out/ncu/step_launches.out 2>&1
tail -2 gpurun_out/ncu/step_launches.out | cut -c1-200 >> $LOG
wc -l gpurun_out/ncu/step_launches.csv >> $LOG
echo "=== ncu --set full: conv / wgrad / BN kernels inside a training step" >> $LOG
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_ws|igemm_persistent|wgrad_tf32|bn_elu_fwd|bn_elu_bwd_reduce|bn_elu_bwd_apply|adam_prox" \
   --launch-skip 330 --launch-count 70 -o gpurun_out/ncu/step_full -f \
   python bench.py --gpus 1 --steps 2 --warmup 3 --no-graphs --no-e2e --no-collective-table > gpurun_out/ncu/step_full.out 2>&1
tail -2 gpurun_out/ncu/step_full.out | cut -c1-200 >> $LOG
echo "=== ncu --set full: aggregation, BB, L-BFGS two-loop, dense, InfoNCE" >> $LOG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"block_reduce|bb_update|lbfgs_two_loop|gemm_f32|info_nce|act_bwd_bias" \
   --launch-count 40 -o gpurun_out/ncu/misc_full -f python tools/ncu_misc.py > gpurun_out/ncu/misc_full.out 2>&1
tail -2 gpurun_out/ncu/misc_full.out | cut -c1-200 >> $LOG
for r in step_full misc_full; do
  if [ -f gpurun_out/ncu/$r.ncu-rep ]; then
    ncu -i gpurun_out/ncu/$r.ncu-rep --page raw --csv > gpurun_out/ncu/${r}_raw.csv 2>/dev/null
    ls -la gpurun_out/ncu/$r.ncu-rep gpurun_out/ncu/${r}_raw.csv >> $LOG
  fi
done
# keep gpurun_out under the 64 MiB merge limit: drop a report that is too large (the CSV stays)
for r in step_full misc_full; do
  f=gpurun_out/ncu/$r.ncu-rep
  if [ -f $f ] && [ $(stat -c %s $f) -gt 25000000 ]; then echo "dropping $f (too large)" >> $LOG; rm -f $f; fi
done
echo "=== compute-sanitizer" >> $LOG
SEL='test_adam_prox_matches_oracle or test_vector_reductions or test_bn_elu_forward_backward or test_cross_entropy_and_vae_loss or (test_conv2d_nhwc_tcgen05 and 4-32-64-64) or (test_conv_kernel_variants and 3-32-4-64) or (test_fused_collective_single_process_matches_torch and 5130) or test_stride2_data_gradient_as_one_stride1_conv or (test_wgrad_matches_fp64_oracle and 5-6-20-12) or (test_wgrad_matches_fp64_oracle and 37-3-64-16) or test_act_bwd_bias or test_maxpool2x2 or test_argmax_count or (test_info_nce_fused and 16-8-3-3) or (test_small_direct_conv and 5-7-9-5) or (test_linear_f32 and 37-3-5) or (test_loopback_fedavg_fedprox_admm and 2-850) or (test_bb_update_kernel and 1)'
for TOOL in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 7 --print-limit 20 \
     python -m pytest tests/test_gpu_kernels.py tests/test_gpu_wgrad.py tests/test_gpu_aux.py tests/test_gpu_loopback.py -m gpu -q -p no:cacheprovider -k "$SEL" > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitize_$TOOL.log
  echo "--- $TOOL" >> $LOG; tail -6 gpurun_out/sanitize_$TOOL.log | cut -c1-300 >> $LOG
done
echo "=== accuracy / rounds-to-target (ResNet18, K = 8 co-resident)" >> $LOG
OUT=gpurun_out/accuracy_gpu NLOOP=2 NEPOCH=8 bash tools/accuracy_gpu.sh >> $LOG 2>&1
echo "=== done" >> $LOG
