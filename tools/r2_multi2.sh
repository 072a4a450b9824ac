#!/bin/bash
# final-state multi-GPU validation: NGPU=2|4|8 bash tools/r2_multi2.sh
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
N=${NGPU:-2}
mkdir -p gpurun_out
LOG=gpurun_out/r2_multi2_n$N.log
: > $LOG
echo "=== cross-rank collective tests (fused vs NCCL)" >> $LOG
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -6 >> $LOG
echo "=== bench N=$N fedavg" >> $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err
tail -3 gpurun_out/final_bench_n$N.err >> $LOG
python - >> $LOG 2>&1 <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/final_bench_n$N.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "aggs", d["config"]["aggregations_in_window"], "e2e", d.get("e2e", {}).get("value"), "clocks", d.get("clocks"))
    c = d.get("collective", {})
    print("collective: world", c.get("world"), "multicast", c.get("multicast_bound"), c.get("transport"), c.get("error"))
    for r in c.get("rows", []):
        print("blk %d n=%8d fused %7.1f us (%s, %6.1f GB/s) p2p %7.1f oneshot %7.1f | nccl %7.1f us allreduce-only %7.1f | x%.2f x%.2f" % (
            r["block"], r["floats"], r["fused_us"], "2shot" if r["fused_two_shot"] else "1shot", r.get("fused_busGBs", 0), r.get("fused_p2p_us", 0),
            r.get("fused_oneshot_us", 0), r.get("nccl_us", 0), r.get("nccl_allreduce_only_us", 0), r.get("speedup_vs_nccl", 0),
            r.get("nccl_allreduce_only_us", 0) / max(r["fused_us"], 1e-9)))
except Exception as e:
    print("parse error", e)
PY
for d in "--driver vae" "--driver fedprox --optimizer lbfgs --no-e2e" "--driver cpc"; do
  echo "=== bench N=$N $d" >> $LOG
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $N --steps 20 --warmup 5 $d --no-collective-table 2>&1 | grep '^{' | tail -1 | cut -c1-2200 >> $LOG
done
echo "=== done" >> $LOG
