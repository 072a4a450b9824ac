#!/bin/bash
# Functional evidence on CPU (no GPU needed): accuracy trajectories of the four CIFAR drivers with the small `Net`,
# K=4 workers, synthetic *learnable* CIFAR10-shaped data (class-conditional templates + noise; not real CIFAR10, so
# the absolute numbers are not comparable with the reference's comparison.png — the shape and the ordering are).
cd "$(dirname "$0")/.."
OUT=profiles/accuracy_cpu
# --data_noise 3.0: noise std = 3x the template amplitude, which puts the small Net in the 40-70 % range (CIFAR10-like difficulty)
COMMON="--K 4 --no-use_cuda --train_size 16000 --test_size 2000 --data_noise 3.0 --no-save_model --check_results"
export OMP_NUM_THREADS=8
python federated_multi.py  $COMMON --Nloop 4 --metrics_path $OUT/fedavg.jsonl   > $OUT/fedavg.log 2>&1
python fedprox_multi.py    $COMMON --Nloop 4 --metrics_path $OUT/fedprox.jsonl  > $OUT/fedprox.log 2>&1
python consensus_multi.py  $COMMON --Nloop 4 --bb_update --metrics_path $OUT/admm_bb.jsonl > $OUT/admm_bb.log 2>&1
python no_consensus_multi.py $COMMON --Nepoch 12 --metrics_path $OUT/no_consensus.jsonl > $OUT/no_consensus.log 2>&1
echo done > $OUT/DONE
