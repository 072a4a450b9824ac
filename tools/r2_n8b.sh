#!/bin/bash
# N=8 re-measurement after moving the nvidia-smi start-up out of the window (+ A/B without the sampler)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
N=${NGPU:-8}
mkdir -p gpurun_out
LOG=gpurun_out/r2_n8b.log
: > $LOG
run() {   # label, env, extra args
  echo "=== $1" >> $LOG
  env $2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $N --steps 20 --warmup 5 $4 > gpurun_out/n8b_$3.json 2> gpurun_out/n8b_$3.err
  python - >> $LOG 2>&1 <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/n8b_$3.json") if l.startswith("{")][-1])
    e = d.get("e2e") or {}
    print("value %.0f ms/step %.3f per_step %s | e2e %s ms/step %s per_step %s" % (d["value"], d["ms_per_step"], d["config"].get("per_step_ms"), e.get("value"), e.get("ms_per_step"), e.get("per_step_ms")))
except Exception as ex:
    print("parse error", ex)
PY
}
run "fedavg (sampler early)" "FEDB200_X=0" 29701 "--no-collective-table"
run "fedavg (no nvidia-smi at all)" "FEDB200_BENCH_NO_SMI=1" 29702 "--no-collective-table"
run "consensus --bb" "FEDB200_X=0" 29703 "--driver consensus --bb --no-e2e --no-collective-table"
echo "=== done" >> $LOG
