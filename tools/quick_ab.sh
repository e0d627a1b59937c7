#!/bin/bash
# A/B of an environment switch on the default bench (run under gpurun): tools/quick_ab.sh VAR tag
VAR=${1:-DFRL_PDL}; TAG=${2:-ab}
for v in 1 0; do
env $VAR=$v timeout 300 python bench.py --no-c5 --no-cpu > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_$v.json')); print('$VAR=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print({k:(v['us_per_launch']) for k,v in d['roofline']['kernels'].items()}); print(d['c2_4096_envs']['ms_per_step'], d['c3_actor_critic_65536_envs']['ms_per_step'])"
done
