"""The HBM-bound kernels of the path timed alone (bench.py's hbm_kernels rows) -- also the command to
profile them:  ncu --set full -k regex:env_step_vec4 python tools/hbm_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dependence_free_rl_b200 as D  # noqa: E402

ctx = D.Context(0)
print(json.dumps(bench.hbm_kernel_rows(D, ctx, bench.load_peaks()), indent=1))
ctx.close()
