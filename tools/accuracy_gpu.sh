#!/bin/bash
# BASELINE metric "rounds-to-target-accuracy" + the ordering of the reference's comparison.png on a B200:
# ResNet18, K = 8 workers with 1/8 of 50 000 synthetic CIFAR10-shaped images each (class templates + noise at 3x the template
# amplitude), 10 000 test images, every worker evaluated after every round (train-mode BatchNorm, Q4), reference defaults
# otherwise.  One GPU, the 8 replicas co-resident on their own streams (gpurun --gpus 1);  NGPU=8 runs one replica per GPU.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=${OUT:-gpurun_out/accuracy_gpu}
mkdir -p $OUT
N=${NGPU:-1}
LAUNCH="python"
if [ "$N" != "1" ]; then LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"; fi
COMMON="--K 8 --use_resnet --data_noise ${NOISE:-3.0} --no-save_model --check_results"
L=${NLOOP:-2}
( time timeout 900 $LAUNCH federated_multi.py  $COMMON --Nloop $L --metrics_path $OUT/fedavg.jsonl ) > $OUT/fedavg.log 2>&1
( time timeout 900 $LAUNCH consensus_multi.py  $COMMON --Nloop $L --bb_update --metrics_path $OUT/admm_bb.jsonl ) > $OUT/admm_bb.log 2>&1
( time timeout 900 $LAUNCH fedprox_multi.py    $COMMON --Nloop 1 --metrics_path $OUT/fedprox.jsonl ) > $OUT/fedprox.log 2>&1
( time timeout 900 $LAUNCH no_consensus_multi.py $COMMON --Nepoch ${NEPOCH:-8} --metrics_path $OUT/standalone_k8.jsonl ) > $OUT/standalone_k8.log 2>&1
( time timeout 900 python no_consensus_multi.py --K 1 --use_resnet --data_noise ${NOISE:-3.0} --no-save_model --check_results --Nepoch ${NEPOCH:-8} --metrics_path $OUT/standalone_k1.jsonl ) > $OUT/standalone_k1.log 2>&1
python tools/accuracy_summary.py $OUT > $OUT/SUMMARY.md 2>&1
cat $OUT/SUMMARY.md
