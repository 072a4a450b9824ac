#!/bin/bash
# Launch list of one eager training step (device time per kernel, serialised) + one full ncu capture of the conv kernel.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
# 1) count launches of the priming+warm-up steps so that we can skip them
SKIP=${SKIP:-4200}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c ${COUNT:-700} --csv --log-file gpurun_out/launches.csv \
   python bench.py --gpus 1 --steps 3 --warmup 3 --no-graphs > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log
# 2) full capture of the hottest kernel (3 instances from the middle of a step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_tf32 -s ${ISKIP:-260} -c 3 -o gpurun_out/prof_igemm -f \
   python bench.py --gpus 1 --steps 2 --warmup 3 --no-graphs > gpurun_out/prof_bench.log 2>&1
tail -2 gpurun_out/prof_bench.log
ls -la gpurun_out
