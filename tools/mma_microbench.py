# Per-instruction pipe cost (SM cycles) of tcgen05.mma forms used by the fused kernels.
import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import dependence_free_rl_b200 as D
ctx = D.Context(0, 1, 0, None)
lib = D._lib.lib
def run(M, N, a, b, n):
    out = (C.c_longlong * 2)()
    D._lib.check(lib.dfrl_umma_microbench(ctx.h, M, N, a, b, n, out))
    return out[0], out[1]
print("M N a_mn b_mn | cycles/MMA (slope 64->256) | fixed | issue cycles/MMA")
for (M, N, a, b) in [(128, 64, 0, 0), (128, 16, 0, 0), (128, 64, 0, 1), (128, 32, 0, 0), (128, 128, 0, 0), (128, 256, 0, 0),
                     (64, 64, 1, 1), (64, 16, 1, 1), (128, 48, 1, 1), (128, 64, 1, 1), (64, 64, 0, 0), (64, 128, 1, 1), (128, 128, 1, 1),
                     (128, 64, 2, 0), (128, 64, 2, 1), (128, 16, 2, 0), (128, 128, 2, 0)]:  # a_mn = 2: A from TMEM
    run(M, N, a, b, 8)
    c1, i1 = run(M, N, a, b, 64)
    c2, i2 = run(M, N, a, b, 256)
    s = (c2 - c1) / 192.0
    print(f"{M:4d} {N:4d} {a} {b} | {s:7.1f} | {c1 - 64 * s:7.0f} | {(i2 - i1) / 192.0:6.1f}")
