#!/bin/bash
# Multi-GPU validation + measurement (gpurun --gpus N): cross-rank kernel tests vs NCCL, bench with the collective table.
#   NGPU=2 bash tools/r2_multi.sh
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
N=${NGPU:-2}
mkdir -p gpurun_out
LOG=gpurun_out/r2_multi_n$N.log
: > $LOG
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
nvidia-smi topo -m 2>&1 | head -14 >> $LOG
echo "=== cross-rank collective tests (fused vs NCCL)" >> $LOG
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider --tb=short -s 2>&1 | tail -40 >> $LOG
echo "=== bench N=$N fedavg" >> $LOG
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err >> $LOG; cut -c1-600 gpurun_out/bench_n$N.json >> $LOG
python - >> $LOG 2>&1 <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "aggs", d["config"]["aggregations_in_window"], "e2e", d.get("e2e", {}).get("value"))
    c = d.get("collective", {})
    print("collective: world", c.get("world"), "multicast", c.get("multicast_bound"), c.get("transport"), c.get("error"))
    for r in c.get("rows", []):
        print("blk %d n=%8d fused %7.1f us (%s, %6.1f GB/s) p2p %7.1f oneshot %7.1f | nccl %7.1f us (%6.1f GB/s) allreduce-only %7.1f | x%.2f" % (
            r["block"], r["floats"], r["fused_us"], "2shot" if r["fused_two_shot"] else "1shot", r.get("fused_busGBs", 0), r.get("fused_p2p_us", 0),
            r.get("fused_oneshot_us", 0), r.get("nccl_us", 0), r.get("nccl_busGBs", 0), r.get("nccl_allreduce_only_us", 0), r.get("speedup_vs_nccl", 0)))
except Exception as e:
    print("parse error", e)
PY
echo "=== bench N=$N consensus --bb" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --driver consensus --bb --no-e2e --no-collective-table 2>&1 | tail -1 | cut -c1-500 >> $LOG
echo "=== bench N=$N --impl nccl (ATen model + NCCL all-reduce: the baseline, not the product)" >> $LOG
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --impl nccl --no-e2e --no-collective-table > gpurun_out/nccl_n$N.out 2>&1
grep -i -m3 "nvls" gpurun_out/nccl_n$N.out | cut -c1-200 >> $LOG
grep '^{' gpurun_out/nccl_n$N.out | tail -1 | cut -c1-400 >> $LOG
echo "=== done" >> $LOG
