#!/bin/bash
# final-state 8-GPU validation and scaling points (N = 8, then N = 4 on the same box)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
LOG=gpurun_out/r2_final_n8.log
: > $LOG
echo "=== cross-rank collective tests (8 ranks)" >> $LOG
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -4 >> $LOG
for N in 8 4; do
  echo "=== bench N=$N fedavg" >> $LOG
  EXTRA=""; if [ $N = 4 ]; then EXTRA="--no-collective-table"; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err
  tail -2 gpurun_out/final_bench_n$N.err | cut -c1-300 >> $LOG
  python - >> $LOG 2>&1 <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/final_bench_n$N.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "aggs", d["config"]["aggregations_in_window"], "e2e", d.get("e2e", {}).get("value"), "clocks", d.get("clocks"), "per_step", d["config"].get("per_step_ms", {}).get("max"))
    c = d.get("collective") or {}
    print("collective: world", c.get("world"), "multicast", c.get("multicast_bound"), c.get("fused_default"), c.get("error"))
    for r in c.get("rows", []):
        print("blk %d n=%8d fused %7.1f us (%s, %6.1f GB/s) p2p %7.1f nvls %7.1f oneshot %7.1f | nccl %7.1f us allreduce-only %7.1f | x%.2f x%.2f" % (
            r["block"], r["floats"], r["fused_us"], "2shot" if r["fused_two_shot"] else "1shot", r.get("fused_busGBs", 0), r.get("fused_p2p_us", 0), r.get("fused_nvls_us", 0),
            r.get("fused_oneshot_us", 0), r.get("nccl_us", 0), r.get("nccl_allreduce_only_us", 0), r.get("speedup_vs_nccl", 0), r.get("speedup_vs_bare_allreduce", 0)))
except Exception as e:
    print("parse error", e)
PY
done
echo "=== bench N=8 consensus --bb" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 --driver consensus --bb --no-e2e --no-collective-table 2>&1 | grep '^{' | tail -1 | cut -c1-330 >> $LOG
echo "=== done" >> $LOG
