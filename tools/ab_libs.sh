#!/bin/bash
# In-run comparison of several builds of the library on ONE GPU box (box-to-box variance is 2-3 %, the
# same box repeats to 0.05 %):   R=2 tools/ab_libs.sh <libA.so> <libB.so> [<libC.so> ...]
# alternates the builds on the default bench R times.
R=${R:-2}
LIB=dependence_free_rl_b200/libdfrl_b200.so
cp $LIB /tmp/lib_keep.so
for r in $(seq 1 $R); do
  for src in "$@"; do
    cp $src $LIB
    python bench.py --no-cpu --no-c5 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('$(basename $src .so)', round(d['ms_per_step'],4), {k.replace('fused_','')[:18]:v['us_per_launch'] for k,v in d['roofline']['kernels'].items()}, 'c2', round(d['c2_4096_envs']['ms_per_step'],4), 'c3', round(d['c3_actor_critic_65536_envs']['ms_per_step'],4))"
  done
done
cp /tmp/lib_keep.so $LIB
