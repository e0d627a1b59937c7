# The reference's own PPO nets (conv1d 4-128-64-1 policy, critic 32-64-32-1) free-running on the fused
# kernels: tools/refnets_run.py [envs] [iterations]   (ncu / timing target)
import sys
sys.path.insert(0, '/root/repo')
import dependence_free_rl_b200 as D
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = D.Context(0, 1, 0, None)
tr, env, pol, val = bench.make_trainer(D, ctx, n, 0, n * 4, player=D.conv_layers([4, 128, 64, 1], D.SOFTMAX), vlayer=D.fc_layers([32, 64, 32, 1]))
tr.iterate(3)
ctx.sync()
ctx.timer_start()
tr.iterate(iters)
ms = ctx.timer_stop() / iters
print(f"{n} envs: {ms:.4f} ms / iteration, {n * 4 / ms / 1e3:.1f} M env-steps/s, coverage {tr.fused_coverage()}")
prof = bench.kernel_profile(D, ctx, tr, 3)   # per-kernel event times (launch by launch, outside the CUDA graph)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k[:70]:70s} {v['launches']:3.0f} launches/iter  {v['ms'] * 1e3 / max(v['launches'], 1):8.1f} us each")
