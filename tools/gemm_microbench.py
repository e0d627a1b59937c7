"""Times the layered dense kernels (dfrl_dense_forward / backward / gradient) for one shape with
CUDA events on the library's stream: python tools/gemm_microbench.py [rows in out]"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import dependence_free_rl_b200 as D

rows, n_in, n_out = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (131072, 256, 256)
ctx = D.Context(0)
lib, chk = D._lib.lib, D._lib.check
rng = np.random.default_rng(0)
p = ctx.to_device((rng.standard_normal((n_in + 1) * n_out) * 0.1).astype(np.float32))
x = ctx.to_device(rng.standard_normal((rows, n_in)).astype(np.float32))
dy = ctx.to_device(rng.standard_normal((rows, n_out)).astype(np.float32))
y = ctx.empty((rows, n_out), np.float32)
gx = ctx.empty((rows, n_in), np.float32)
g = ctx.empty(((n_in + 1) * n_out,), np.float32)
flops = 2.0 * rows * n_in * n_out
def t(f, n=10):
    for _ in range(3): f()
    ctx.sync(); ctx.timer_start()
    for _ in range(n): f()
    return ctx.timer_stop() / n
ms = t(lambda: chk(lib.dfrl_dense_forward(ctx.h, p.p, n_in, n_out, x.p, rows, y.p, 1)))
print(f"forward  {rows}x{n_in}->{n_out}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s")
ms = t(lambda: chk(lib.dfrl_dense_backward(ctx.h, p.p, n_in, n_out, dy.p, rows, x.p, gx.p)))
print(f"backward {rows}x{n_out}->{n_in}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s")
ms = t(lambda: chk(lib.dfrl_dense_gradient(ctx.h, n_in, n_out, x.p, dy.p, rows, g.p, 0)))
print(f"gradient {rows}: dW {n_out}x{n_in}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s")
