# Phase clocks (SM cycles) of warpgroup 0 of CTA 0 of fused_policy_step_kernel: 13 stamps per tile.
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
import dependence_free_rl_b200 as D
import bench
ctx = D.Context(0, 1, 0, None)
tr, env, pol, val = bench.make_trainer(D, ctx, 131072, 0, 131072*4)
lib = D._lib.lib
buf = (C.c_longlong * 112)()
D._lib.check(lib.dfrl_debug_policy_clocks(tr.h, buf, 112))
tr.iterate(3)
D._lib.check(lib.dfrl_debug_policy_clocks(tr.h, buf, 112))
a = np.array(buf[:104]).reshape(8, 13)
names = ["L1 mma wait", "epi1+sync", "L2 wait", "epi2+sync", "L3 wait", "head+sync", "dW3+dH2 wait", "epi dH2+sync",
         "dH1 wait+slot", "epi dH1+sync", "dW1 (+dW2) wait", "encode X0+sync"]
d = np.diff(a, axis=1)
print("tile totals of one warpgroup (cycles):", (a[1:, 0] - a[:-1, 0]))
for j, nm in enumerate(names):
    print(f"{nm:18s}", d[1:, j])
e = np.array(buf[104:110])
print("entry -> setup done", e[1] - e[0], " setup -> loop end", e[2] - e[1], " drain (TMEM -> partials)", e[4] - e[2],
      " grid barrier", e[5] - e[4], " slice reduction + update", e[3] - e[5], " total", e[3] - e[0])
