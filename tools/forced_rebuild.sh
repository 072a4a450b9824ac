#!/bin/bash
# Forced from-scratch build of every native extension through __graft_entry__.build()'s code path (torch.utils.cpp_extension with
# explicit -gencode arch=compute_100a,code=sm_100a), logging the nvcc / g++ command lines and the ptxas -v resource summary.
#   bash tools/forced_rebuild.sh  -> profiles/r2/build_forced_rebuild.log       (runs WITHOUT a GPU)
cd "$(dirname "$0")/.."
LOG=profiles/r2/build_forced_rebuild.log
rm -rf federated_pytorch_test_b200/_build
( time python -c "
import sys; sys.path.insert(0, '.')
from federated_pytorch_test_b200 import _ext
_ext.build_all(verbose=True)
print('BUILD OK', [_ext._so_path(n) for n in _ext._SPECS])
" ) > /tmp/_build_full.log 2>&1
{
  echo "# Forced rebuild of every native extension from __graft_entry__.build()'s code path (torch.utils.cpp_extension, explicit flags), CPU box, $(date -u +%F)"
  grep -E "^\[[0-9]+/[0-9]+\]" /tmp/_build_full.log | cut -c1-1600
  echo
  echo "# ptxas -v: kernels compiled / stack frames and spills (count, text)"
  grep -c "Compiling entry function" /tmp/_build_full.log
  grep -E "bytes stack frame" /tmp/_build_full.log | sed 's/^ *//' | sort | uniq -c | sort -rn | head -8
  echo "# ptxas -v: highest register counts"
  grep -E "Used [0-9]+ registers" /tmp/_build_full.log | sed -E 's/.*Used ([0-9]+) registers.*/\1/' | sort -n | uniq -c | tail -5
  grep -E "BUILD OK|^real|rror" /tmp/_build_full.log | cut -c1-400
} > $LOG 2>&1
python __graft_entry__.py >> $LOG 2>&1
tail -4 $LOG
