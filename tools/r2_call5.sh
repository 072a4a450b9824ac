#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_call5.log
: > $LOG
echo "=== loopback + aux + kernels tests" >> $LOG
timeout 900 python -m pytest tests/test_gpu_loopback.py tests/test_gpu_aux.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 >> $LOG
echo "=== wgrad micro-benchmark" >> $LOG
timeout 600 python tools/bench_wgrad.py >> $LOG 2>&1
echo "=== wgrad gpc sweep (layer1, layer2, layer4)" >> $LOG
for g in 1 2 3; do echo "gpc=$g" >> $LOG; FEDB200_WGRAD_GPC=$g timeout 300 python tools/bench_wgrad.py layer >> $LOG 2>&1; done
echo "=== ResNet18 whole-network gradient error budget" >> $LOG
timeout 600 python tools/measure_resnet_err.py >> $LOG 2>&1
echo "=== aggregation kernel K=8 co-resident (1 GPU), coop vs plain launch" >> $LOG
timeout 300 python - >> $LOG 2>&1 <<'PY'
import torch, os, sys
sys.path.insert(0, ".")
from federated_pytorch_test_b200.parallel import Topology
from federated_pytorch_test_b200.parallel.fused import FusedCollective
dev = torch.device("cuda", 0)
for K in (1, 8):
    topo = Topology.single_process(K, dev)
    coll = FusedCollective(topo)
    for n in (1856, 73984, 1180672, 4720640):
        st = -(-n // 32) * 32
        arena = coll.heap.alloc(K * st)
        xs = [arena[k * st: k * st + n] for k in range(K)]
        z = coll.zeros_like_block(xs[0], "z") if K == 1 else torch.zeros(n, device=dev)
        for _ in range(3):
            coll._launch(0, xs, None, z, 0.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            coll._launch(0, xs, None, z, 0.0)
        b.record(); torch.cuda.synchronize()
        print("K=%d n=%8d fedavg %.1f us" % (K, n, a.elapsed_time(b) * 1e3 / 20), flush=True)
PY
echo "=== bench x2" >> $LOG
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-collective-table 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['config']['per_step_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" >> $LOG 2>&1
done
echo "=== aux benches (vae, vae_cl, cpc) N=1" >> $LOG
for d in vae vae_cl cpc; do timeout 300 python bench.py --driver $d --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-330 >> $LOG; done
echo "=== reference vae" >> $LOG
timeout 600 python bench.py --impl reference --driver vae --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-330 >> $LOG
echo "=== done" >> $LOG
