# Phase clocks (SM cycles) of warpgroup 0 of CTA 0 of fused_policy_step_kernel: 13 stamps per tile.
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
import dependence_free_rl_b200 as D
import bench
ctx = D.Context(0, 1, 0, None)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
tr, env, pol, val = bench.make_trainer(D, ctx, N, 0, N*4)
lib = D._lib.lib
buf = (C.c_longlong * 752)()
D._lib.check(lib.dfrl_debug_policy_clocks(tr.h, buf, 112))
tr.iterate(3)
D._lib.check(lib.dfrl_debug_policy_clocks(tr.h, buf, 752))
a = np.array(buf[:104]).reshape(8, 13)
if N * 4 // 128 <= 148:   # one tile per CTA: the stamps of the single tile only
    print("single tile (cycles):", dict(zip(["L1 mma wait", "epi1+sync", "L2 wait", "epi2+sync", "L3 wait", "head+sync", "dW3+dH2 wait",
          "epi dH2+sync", "dH1 wait+slot", "epi dH1+sync", "dW1 (+dW2) wait", "encode X0+sync"], np.diff(a[0]).tolist())))
names = ["L1 mma wait", "epi1+sync", "L2 wait", "epi2+sync", "L3 wait", "head+sync", "dW3+dH2 wait", "epi dH2+sync",
         "dH1 wait+slot", "epi dH1+sync", "dW1 (+dW2) wait", "encode X0+sync"]
d = np.diff(a, axis=1)
import os
cta = int(os.environ.get("DFRL_CLK_CTA", "0"))
print(f"phase clocks of CTA {cta}")
print("tile totals of one warpgroup (cycles):", (a[1:, 0] - a[:-1, 0]))
for j, nm in enumerate(names):
    print(f"{nm:18s}", d[1:, j])
e = np.array(buf[104:112])
print("entry -> setup done", e[1] - e[0], " setup -> loop end", e[2] - e[1], " drain (TMEM -> partials)", e[4] - e[2],
      " grid barrier", e[5] - e[4], " partial loads + sum", e[6] - e[5], " exchange + update + exit", e[3] - e[6], " total", e[3] - e[0])

# every CTA of the last launch: wall clock (ns, %globaltimer) at entry / tile-loop end / kernel end, and its SM
g = np.array(buf[112:112 + 4 * 148]).reshape(148, 4)
t0 = g[:, 0].min()
ent, loop, end = g[:, 0] - t0, g[:, 1] - g[:, 0], g[:, 2] - g[:, 1]
print("per-CTA (ns): entry skew min/med/max", ent.min(), int(np.median(ent)), ent.max(), " tile loop min/med/max", loop.min(), int(np.median(loop)), loop.max(),
      " loop end (abs) min/med/max", (g[:, 1] - t0).min(), int(np.median(g[:, 1] - t0)), (g[:, 1] - t0).max(), " drain+tail min/med/max", end.min(), int(np.median(end)), end.max())
nt = np.array([(N * 4 // 128 - 1 - b) // 148 + 1 if b < N * 4 // 128 else 0 for b in range(148)])
for k in sorted(set(nt)):
    m = nt == k
    print(f"  CTAs with {k} tiles: {m.sum():3d}  loop ns min/med/max {loop[m].min()} {int(np.median(loop[m]))} {loop[m].max()}")
slow = np.argsort(-(g[:, 1] - t0))[:8]
print(f"  CTA {cta} ran on SM {int(g[cta, 3])}: {loop[cta] / max(nt[cta], 1):.0f} ns / tile")
print("  latest loop ends: CTA, SM, tiles, entry, loop:", [(int(b), int(g[b, 3]), int(nt[b]), int(ent[b]), int(loop[b])) for b in slow])
if len(sys.argv) > 2:   # full table: per-tile loop time by SM
    order = np.argsort(g[:, 3])
    print("SM:ns/tile", " ".join(f"{int(g[b, 3])}:{loop[b] / max(nt[b], 1):.0f}" for b in order))
if len(sys.argv) > 2:   # ... and by CTA index
    print("CTA:SM:ns/tile", " ".join(f"{b}:{int(g[b, 3])}:{loop[b] / max(nt[b], 1):.0f}" for b in range(148)))
