#!/bin/bash
# Multi-GPU check: fused collectives across ranks (P2P + multimem) and the scaling bench at N GPUs.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
N=${NGPU:-2}
OUT=gpurun_out/multi_$N.log
: > $OUT
nvidia-smi topo -m >> $OUT 2>&1
echo "=== test_gpu_multi" >> $OUT
NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s -p no:cacheprovider 2>&1 | tail -30 >> $OUT
echo "=== bench ours N=$N" >> $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps ${STEPS:-60} --warmup 5 2>&1 | tail -8 >> $OUT
echo "=== bench nccl baseline N=$N" >> $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --impl nccl --gpus $N --steps ${STEPS:-60} --warmup 5 2>&1 | tail -5 >> $OUT
if [ -n "$WITH_REF" ]; then
echo "=== bench reference N=$N" >> $OUT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus $N --steps ${STEPS:-60} --warmup 5 2>&1 | tail -5 >> $OUT
fi
echo "=== done" >> $OUT
tail -3 $OUT
