#!/bin/bash
# In-run comparison of library builds on the reference-nets row (table path, 131 072 envs):
#   R=2 tools/ab_refnets.sh <libA.so> <libB.so> ...     (run under gpurun; see ab_libs.sh)
R=${R:-2}
LIB=dependence_free_rl_b200/libdfrl_b200.so
cp $LIB /tmp/lib_keep.so
for r in $(seq 1 $R); do
  for src in "$@"; do
    cp $src $LIB
    echo "== $(basename $src .so)"
    python tools/refnets_run.py 131072 200 2>/dev/null | grep -E "envs:|head_kernel|backward_kernel|forward_kernel|rollout_kernel"
  done
done
cp /tmp/lib_keep.so $LIB
