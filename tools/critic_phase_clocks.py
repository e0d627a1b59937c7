# Phase clocks (SM cycles) of pipeline 0 of CTA 0 of fused_critic_kernel<CRITIC_STEP>: 15 stamps per tile.
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
import dependence_free_rl_b200 as D
import bench
ctx = D.Context(0, 1, 0, None)
tr, env, pol, val = bench.make_trainer(D, ctx, 131072, 0, 131072*4)
lib = D._lib.lib
buf = (C.c_longlong * 752)()
D._lib.check(lib.dfrl_debug_critic_clocks(tr.h, buf, 112))
tr.rollout(); tr.learn(); tr.rollout(); tr.learn()   # plain launches (a captured graph keeps its old arguments)
D._lib.check(lib.dfrl_debug_critic_clocks(tr.h, buf, 752))
a = np.array(buf[:105]).reshape(7, 15)
names = ["L1e wait", "epi1e", "dW1 wait + XS encode", "loads + L2e wait", "pack + value_e", "L1s wait", "epi1s",
         "end-state pack", "L2s wait", "value_s + exchange", "targets + dH2", "dH1 wait + slot", "epi dH1",
         "dW2 wait + staging"]
d = np.diff(a, axis=1)
print("tile totals of one pipeline (cycles):", (a[1:, 0] - a[:-1, 0]))
for j, nm in enumerate(names):
    print(f"{nm:24s}", d[1:, j])

# every CTA of the last critic-step launch: wall clock (ns) at entry / tile-loop end / kernel end, and its SM
g = np.array(buf[112:112 + 4 * 148]).reshape(148, 4)
loop, end = g[:, 1] - g[:, 0], g[:, 2] - g[:, 1]
print("per-CTA (ns): tile loop min/med/max", loop.min(), int(np.median(loop)), loop.max(), " drain+tail min/med/max", end.min(), int(np.median(end)), end.max())
