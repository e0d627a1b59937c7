#!/bin/bash
# One-GPU round-end measurement (run under gpurun from the repo root): GPU test suite, smoke, the
# default bench, phase clocks of the policy / critic kernels, the ncu launch list and ONE full ncu
# capture of the fused kernels. Everything lands in gpurun_out/ with the given tag.
TAG=${1:-rXX}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 300 gpurun_out/bench_${TAG}.err
timeout 100 python tools/policy_phase_clocks.py > gpurun_out/clocks_policy_${TAG}.txt 2>&1
timeout 100 python tools/critic_phase_clocks.py > gpurun_out/clocks_critic_${TAG}.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-c2 --no-c5 > gpurun_out/ncu_l.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fused_ -s 24 -c 6 -f \
  -o gpurun_out/prof_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu --no-c2 --no-c5 > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log | cut -c1-80
