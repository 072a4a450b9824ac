#!/bin/bash
# What round 2 left UNMEASURED (its GPU budget ended on an 8-GPU call, profiles/r2_scaling.md).  Run first in a next round:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_validate.sh'                 (1 GPU, ~3 min)
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'NGPU=8 bash tools/next_round_validate.sh'  (8 GPUs, ~2 min: the open question)
# Open question: with deferred rounds (algo/engine.py::_aggregate) the FIRST aggregation of an 8-GPU run cost 11 ms once and an adaptive-ADMM run
# was slow (5.1 ms/step) while 1 / 4 GPUs and the second aggregation at 8 GPUs were fine.  Since then: deferral limited to <= 4 ranks, the pinned
# record copy only on the deferred path, bench window moved to the second round boundary.  The A/B below decides whether deferral can be enabled at 8.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
N=${NGPU:-1}
mkdir -p gpurun_out
LOG=gpurun_out/next_round_validate_n$N.log
: > $LOG
run() {   # run <label> <env assignments...> -- bench args
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $label" >> $LOG
  if [ $N = 1 ]; then
    env "${envs[@]}" timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-collective-table "$@" 2>&1 | grep '^{' | tail -1 | python tools/bench_brief.py >> $LOG
  else
    env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
        bench.py --gpus $N --steps 20 --warmup 5 --no-collective-table "$@" 2>&1 | grep '^{' | tail -1 | python tools/bench_brief.py >> $LOG
  fi
}
if [ $N = 1 ]; then
  echo "=== full GPU suite" >> $LOG
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -4 >> $LOG
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $LOG
fi
run "fedavg, defaults (deferral: on up to 4 ranks)" X=1 --
run "fedavg, deferred rounds forced ON" FEDB200_DEFERRED_ROUNDS=1 --
run "fedavg, deferred rounds OFF" FEDB200_DEFERRED_ROUNDS=0 --
run "adaptive ADMM, deferred forced ON" FEDB200_DEFERRED_ROUNDS=1 -- --driver consensus --bb --no-e2e
run "adaptive ADMM, deferred OFF" FEDB200_DEFERRED_ROUNDS=0 -- --driver consensus --bb --no-e2e
cat $LOG
