# Phase clocks (SM cycles) of one epilogue thread of CTA 0 of fused_conv_policy_step_kernel: 7 stamps per tile.
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
import dependence_free_rl_b200 as D
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
ctx = D.Context(0, 1, 0, None)
tr, env, pol, val = bench.make_trainer(D, ctx, n, 0, n * 4, player=D.conv_layers([4, 128, 64, 1], D.SOFTMAX), vlayer=D.fc_layers([32, 64, 32, 1]))
lib = D._lib.lib
buf = (C.c_longlong * 112)()
D._lib.check(lib.dfrl_debug_policy_clocks(tr.h, buf, 112))
tr.iterate(3)
D._lib.check(lib.dfrl_debug_policy_clocks(tr.h, buf, 112))
a = np.array(buf[:98]).reshape(14, 7)
names = ["L1 wait (+dW2 of prev)", "epi1", "L2 wait", "head + dH2", "dH1 wait (+dW1 of prev)", "epi dH1 + stage X"]
d = np.diff(a, axis=1)
print("tile totals (cycles):", a[1:, 0] - a[:-1, 0])
for j, nm in enumerate(names):
    print(f"{nm:26s}", d[1:, j])
