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
a = np.array(buf[:112]).reshape(8, 14)
names = ["(unused)","L1 mma wait","epi1+sync","L2 wait","epi2+sync","L3 wait","wg1 stash+bar","wg1 encode","wait for wg0 head+sync","dH2 wait","epi dH2+sync","dH1 wait","epi dH1+sync"]
d = np.diff(a, axis=1)
print("tile totals (cycles):", (a[1:,0]-a[:-1,0]))
for j,nm in enumerate(names):
    print(f"{nm:16s}", d[1:,j])
print("gap tile end -> next tile start", a[1:,0]-a[:-1,13])
