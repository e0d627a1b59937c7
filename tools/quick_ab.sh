for coop in 1 0; do
DFRL_COOP=$coop timeout 300 python bench.py --no-c5 --no-cpu > gpurun_out/r02_bench_coop$coop.json 2> gpurun_out/r02_bench_coop$coop.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_coop$coop.json')); print('coop=$coop', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print({k:(v['us_per_launch']) for k,v in d['roofline']['kernels'].items()}); print(d['c2_4096_envs']['ms_per_step'], d['c3_actor_critic_65536_envs']['ms_per_step'])"
done
