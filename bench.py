#!/usr/bin/env python
"""bench.py -- PPO bin-packing env-steps/s on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (libdfrl_b200.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU trainer

A "step" is one PPO iteration of the reference trainer main (ppo_training.cc:53-66): a rollout
of T = 4 steps of every environment, then learner.step() (1 critic update + k = 4 policy updates
on all N*T transitions), then forget().  Environments are sharded across ranks; parameters are
replicated and each optimizer step all-reduces (SUM) the flat gradient.

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 4            # ppo_training.cc:31 steps_per_worker
EPOCHS = 4             # policy_gradient.h:300
POLICY_DIMS = [32, 64, 64, 8]   # C2/C4: "2-layer 64-hidden MLP", B = 8 bins
VALUE_DIMS = [32, 64, 64, 1]
REF_LR_P, REF_LR_V, REF_ROWS = 1e-4, 1e-5, 32  # ppo_training.cc:17,26; 8 workers x 4 steps


def flops_per_env_step(pd=POLICY_DIMS, vd=VALUE_DIMS, T=T_STEPS, k=EPOCHS):
    """Algorithmic FLOPs of the reference schedule per transition:
    F_pi (1 + 3k (T+1)/T) + F_V 5 (T+1)/T with F = sum 2 in out."""
    fp = sum(2 * a * b for a, b in zip(pd[:-1], pd[1:]))
    fv = sum(2 * a * b for a, b in zip(vd[:-1], vd[1:]))
    return fp * (1 + 3 * k * (T + 1) / T) + fv * 5 * (T + 1) / T


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.stop, self.th = device, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------ reference / CPU arm ------

def cpu_reference_rate(n_envs, iters, threads, seed=1234):
    """Times the reference's own PPO trainer (oracle/_ref, unmodified reference sources) with the
    bench config's nets on the host cores: rollouts on `threads` worker threads exactly as
    ppo_training.cc does, learner single-threaded (the reference has no parallel learner).
    Falls back to the plain-C oracle port when the compiled reference is not present."""
    from oracle import ref as R
    if R.available():
        pol = R.fc_net(POLICY_DIMS, R.SOFTMAX)
        val = R.fc_net(VALUE_DIMS)
        pp, vp = R.init_params(pol, seed), R.init_params(val, seed + 1)
        rows = n_envs * T_STEPS
        res = R.train(R.PPO, seed, n_envs, T_STEPS, iters, pol, pp, REF_LR_P * REF_ROWS / rows, val, vp,
                      REF_LR_V * REF_ROWS / rows, record=False, threads=threads)
        return res["env_steps"] / res["seconds"], res["seconds"], "reference", threads
    from oracle import orc
    ecfg = orc.env_cfg(8)
    pnet, vnet = orc.fc_net(POLICY_DIMS, orc.SOFTMAX), orc.fc_net(VALUE_DIMS)
    rng = np.random.default_rng(seed)
    pp = (rng.standard_normal(pnet.param_count()) * 0.01).astype(np.float32)
    vp = (rng.standard_normal(vnet.param_count()) * 0.01).astype(np.float32)
    rows = n_envs * T_STEPS
    lr = orc.Learner(orc.train_cfg(orc.PPO, T_STEPS, policy_lr=REF_LR_P * REF_ROWS / rows,
                                   value_lr=REF_LR_V * REF_ROWS / rows), ecfg, pnet, pp, vnet, vp)
    st = orc.env_reset_all(ecfg, n_envs, rng.integers(0, 2, n_envs).astype(np.uint8))
    t0 = time.perf_counter()
    for _ in range(iters):
        items = rng.integers(0, 2, (T_STEPS, n_envs)).astype(np.uint8)
        ro = orc.rollout(ecfg, st, pnet, lr.pparams, T_STEPS, 0, items, u=rng.random((T_STEPS, n_envs)))
        lr.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
    dt = time.perf_counter() - t0
    return n_envs * T_STEPS * iters / dt, dt, "port", 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # Bounded sample: the reference's cost is flat per transition (~0.45 ms), so
    # the envs per step are sized for about one minute of CPU work over the whole W + K run.
    n_envs = args.ref_envs
    if n_envs <= 0:
        n_envs = int(max(16, min(256, 60.0 * 2200.0 / (T_STEPS * (args.steps + max(1, args.warmup))))))
    # warm-up then K steps; one step = one PPO iteration on the bounded sample of n_envs envs
    cpu_reference_rate(n_envs, max(1, args.warmup), cores)
    rate, secs, kind, used = cpu_reference_rate(n_envs, args.steps, cores)
    line = {
        "impl": "reference", "metric": "ppo_binpacking_env_steps_per_sec", "value": rate,
        "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_envs_note=f"bounded CPU sample: {n_envs} envs x T={T_STEPS} per step"),
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": used, "kind": kind,
                         "sample": f"{args.steps} PPO iterations of {n_envs} envs x {T_STEPS} steps, reference "
                                   f"nets of the bench config, rollouts on {used} threads, learner 1 thread"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_envs_note=None):
    return {
        "workload": (f"PPO-clip bin packing (8 bins), {args.envs_per_gpu} envs/GPU x {args.gpus} GPU, "
                     f"T={T_STEPS} steps/iter, k={EPOCHS} epochs, policy 32-64-64-8 softmax, value 32-64-64-1 "
                     f"(BASELINE configs[3] 'PPO, 1M envs sharded over 8 B200' at 131072 envs/GPU; "
                     f"same nets/schedule as configs[1])"),
        "envs_per_gpu": args.envs_per_gpu, "global_envs": args.envs_per_gpu * args.gpus,
        "steps_per_iter": T_STEPS, "epochs": EPOCHS, "optimizer": "sgd",
        "lr": "reference 1e-4 / 1e-5 scaled by 32 / (global rows) because gradients are SUMS over rows",
        "items": "synthetic Bernoulli(0.4) stream, Philox4x32-10 keyed (seed, global env, draw)",
        "timing": "value = median of R blocks of exactly --steps iterations (R = repeats: >= --min-timed-s of device time), "
                  "each block between barrier + synchronize, CUDA events, max over ranks; e2e the same with the wall clock",
        "parallelism": f"dp{args.gpus} (envs sharded, flat-gradient all-reduce SUM)",
        "l2": "activation working set exceeds the 126 MB L2 at 131072 envs/GPU; no flush needed",
        **({"note": n_envs_note} if n_envs_note else {}),
    }


# ------------------------------------------------------------------ our arm ------------------

C5_POLICY_DIMS = [128, 256, 256, 256, 32]   # BASELINE configs[4]: 32 bins, 3 hidden layers of 256
C5_VALUE_DIMS = [128, 256, 256, 256, 1]


def make_trainer(D, ctx, n_envs, env_offset, global_rows, seed=1234, fused=1, pdims=None, vdims=None, n_bins=8,
                 algo=None, work=T_STEPS, last=None, opt=None, shared=False, player=None, vlayer=None, lr_scale=1.0):
    pdims, vdims = pdims or POLICY_DIMS, vdims or VALUE_DIMS
    policy = D.Model(ctx, player or D.fc_layers(pdims, last if last is not None else D.SOFTMAX), 4 * n_bins)
    policy.init_parameters(seed)       # identical on every rank (replicated parameters)
    if shared:   # BASELINE configs[2]: one trunk 32-64-64, heads 64-8 and 64-1 (dfrl_mlp_create_shared)
        value = D.Model.shared(policy, 4, [(D.DENSE, pdims[-2], 1)])
    else:
        value = D.Model(ctx, vlayer or D.fc_layers(vdims), 4 * n_bins)
    value.init_parameters(seed + 1)
    env = D.Environment(ctx, n_envs, n_bins=n_bins, seed=seed, env_offset=env_offset)
    kw = {} if opt is None else {"policy_opt": opt, "value_opt": opt}
    tr = D.Trainer(ctx, env, policy, value, algo=algo if algo is not None else D.PPO, work=work,
                   policy_lr=lr_scale * REF_LR_P * REF_ROWS / global_rows, value_lr=lr_scale * REF_LR_V * REF_ROWS / global_rows,
                   fused=fused, **kw)
    return tr, env, policy, value


def quick_rate(D, ctx, objs, n_envs, work, warm, iters):
    """env-steps/s of `iters` device-resident iterations (CUDA events), then closes the objects."""
    tr = objs[0]
    tr.iterate(warm)
    ctx.sync()
    l0 = ctx.launches()
    ctx.timer_start()
    tr.iterate(iters)
    ms = ctx.timer_stop() / iters
    launches = (ctx.launches() - l0) / iters
    for o in (objs[0], objs[1], objs[3], objs[2]):   # trainer, env, value (a sharer goes first), policy
        o.close()
    return {"value": n_envs * work / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "envs": n_envs,
            "steps_per_iter": work, "launches_per_step": launches}


def hbm_kernel_rows(D, ctx, peaks):
    """The HBM-bound kernels north_star names (env step, GAE, optimizer, heuristic play), each timed
    alone on L2-exceeding inputs: achieved GB/s = SURVEY section 8d's algorithmic bytes / time."""
    lib, chk = D._lib.lib, D._lib.check
    rows = {}

    def timed(fn, reps):
        fn()
        ctx.sync()
        ctx.timer_start()
        for _ in range(reps):
            fn()
        return ctx.timer_stop() / reps

    def row(name, ms, nbytes, what):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows[name] = {"us_per_launch": round(1e3 * ms, 2), "algorithmic_mb_per_launch": round(nbytes / 1e6, 2),
                      "gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4), "what": what}
    # K1 environment::apply + game_over + reset + next item, 8 Mi envs (151 MB of state: > L2)
    n = 1 << 23
    env = D.Environment(ctx, n, seed=5)
    act = ctx.to_device(np.random.default_rng(0).integers(0, 8, n).astype(np.uint8))
    done = ctx.empty((n,), np.uint8)
    ms = timed(lambda: chk(lib.dfrl_env_step(env.h, act.p, done.p, None)), 20)
    row("env_step_vec4_kernel", ms, n * 38, "8 Mi envs, 8 bins: 2 x 18 B state + action + done per env-step (SURVEY 8d K1)")
    # batched heuristic policy + env, whole episodes on the device (min-waste, 1 episode per env)
    tot, steps = C.c_double(), C.c_longlong()
    env.reset()
    ctx.sync()
    ctx.timer_start()
    chk(lib.dfrl_heuristic_play(env.h, D.HEUR_MINWASTE, 1, C.byref(tot), C.byref(steps)))
    ms = ctx.timer_stop()
    rows["heuristic_play_kernel"] = {"ms": round(ms, 3), "env_steps": steps.value, "env_steps_per_s": steps.value / (ms * 1e-3),
                                      "mean_reward": tot.value / n,
                                      "what": "min-waste policy, 8 Mi envs x 1 episode, state resident in L1/L2 per thread: "
                                              "latency / instruction bound, not an HBM stream"}
    for o in (act, done):
        o.free()
    env.close()
    # K4 GAE on [T][n] records, T = 4, 16 Mi envs, 8 % of the steps end an episode (random-policy level)
    T, n = 4, 1 << 24
    dh = (np.random.default_rng(1).random((T, n), dtype=np.float32) < 0.08).astype(np.uint8)
    ends = int(dh[:T - 1].sum()) + n           # rows whose V(next) comes from v_end: done, or the last step
    d = ctx.to_device(dh)
    vs, ve = ctx.zeros((T, n), np.float32), ctx.zeros((T, n), np.float32)
    tg, adv = ctx.empty((T, n), np.float32), ctx.empty((T, n), np.float32)
    ms = timed(lambda: chk(lib.dfrl_gae(ctx.h, d.p, vs.p, ve.p, n, T, C.c_float(0.99), C.c_float(0.95), tg.p, adv.p)), 10)
    row("gae_kernel", ms, T * n * 13 + 4 * ends,
        "16 Mi envs x T=4: done + V_t read, target + advantage written = 13 B/row, + 4 B of V(end) for the "
        f"{100.0 * ends / (T * n):.0f} % of rows that end a trajectory (SURVEY 8d K4 counts 17 B/row: V_t+1 re-read)")
    for o in (d, vs, ve, tg, adv):
        o.free()
    # K7 optimizers, 64 Mi parameters
    P = 1 << 26
    prm, g, st = ctx.zeros((P,), np.float32), ctx.zeros((P,), np.float32), ctx.zeros((2 * P,), np.float32)
    for kind, name, bpp in ((D.SGD, "opt_kernel<sgd>", 12), (D.ADAM, "opt_kernel<adam>", 28)):
        ms = timed(lambda: chk(lib.dfrl_opt_step(ctx.h, kind, prm.p, g.p, st.p, P, C.c_float(1e-4), C.c_float(0.0),
                                                 C.c_float(0.9), C.c_float(0.999), C.c_float(3.0))), 10)
        row(name, ms, P * bpp, f"64 Mi parameters, {bpp} B/param (K7)")
    for o in (prm, g, st):
        o.free()
    return rows


def cpu_reference_rows(cores):
    """CPU rows of SURVEY section 8d on the box's host cores (compiled reference, oracle/_ref):
    C1 = REINFORCE, single env, reference default FC policy 32-256-128-8 (BASELINE configs[0]);
    the reference's own PPO nets (conv1d 4-128-64-1 + FC 32-64-32-1) at 8 workers x 4 steps."""
    from oracle import ref as R
    if not R.available():
        return None
    out = {}
    pol = R.fc_net([32, 256, 128, 8], R.SOFTMAX_CE)
    res = R.train(R.REINFORCE, 1234, 1, 1, 300, pol, R.init_params(pol, 1), 1e-4, record=False, threads=1)
    out["c1_reinforce_single_env_cpu"] = {"value": res["env_steps"] / res["seconds"], "unit": "env-steps/s", "cores": 1,
                                          "sample": f"300 updates of 1 env x 1 episode ({res['seconds']:.1f} s), pg_training.cc nets"}
    pol, val = R.conv_net([4, 128, 64, 1], R.SOFTMAX), R.fc_net([32, 64, 32, 1])
    res = R.train(R.PPO, 1234, 8, 4, 60, pol, R.init_params(pol, 1), 1e-4, val, R.init_params(val, 2), 1e-5,
                  record=False, threads=min(8, cores))
    out["ppo_reference_nets_cpu"] = {"value": res["env_steps"] / res["seconds"], "unit": "env-steps/s", "cores": min(8, cores),
                                     "sample": f"60 PPO rounds of 8 workers x 4 steps ({res['seconds']:.1f} s), ppo_training.cc nets and rates"}
    return out


def measure(D, ctx, dist, args, n_envs, world, rank, steps, warmup, sample_clocks):
    """Returns dict(value, ms_per_step, e2e, launches, clocks) for n_envs envs per rank."""
    global_rows = n_envs * world * T_STEPS
    tr, env, policy, value = make_trainer(D, ctx, n_envs, rank * n_envs, global_rows, fused=args.fused)
    lib = D._lib.lib

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- device-resident throughput: `steps` iterations, no host round trips inside.
    # The clock sampler (nvidia-smi every 50 ms) runs from the warm-up to the end of the e2e
    # measurement: one step is ~1 ms, so the timed region alone is shorter than a sampling period
    # unless --steps is in the hundreds.
    sampler = ClockSampler(ctx_device(ctx)) if sample_clocks else None
    if sampler:
        sampler.__enter__()
    tr.iterate(warmup)

    def timed_block():
        """EXACTLY `steps` iterations between a barrier + synchronize on both sides, device time (CUDA
        events on the library's stream), max over ranks."""
        barrier()
        ctx.timer_start()
        tr.iterate(steps)
        ms = ctx.timer_stop()
        barrier()
        return max_over_ranks(ms)
    l0 = ctx.launches()
    first = timed_block()
    launches = ctx.launches() - l0
    # the block is repeated until >= min_timed_s of device time has been measured (a 20-step block is
    # 12 ms: too short for the clock sampler and for the power state to settle); the MEDIAN block is
    # reported. Every rank derives the same repeat count from the max-over-ranks first block.
    repeats = int(min(400, max(1, np.ceil(args.min_timed_s * 1e3 / max(first, 1e-3)))))
    blocks = [first] + [timed_block() for _ in range(repeats - 1)]
    ms = float(np.median(blocks))
    value_rate = n_envs * world * T_STEPS * steps / (ms * 1e-3)

    # ---- e2e through the public call with HOST buffers: per step the item stream of the step
    # (uint8 [T][n], pinned) goes host->device inside rollout(), learn() runs, and the step's
    # result (env-steps / episodes / reward counters) comes back device->host.
    # The loop keeps one step in flight, as a training loop that logs statistics would: step i's
    # result is read (stats_begin: async D2H + event) and waited for only after step i + 1 has been
    # submitted, and the host fills the OTHER of two pinned tapes meanwhile. Every step's tape copy
    # and result read happen inside the timed region.
    nbytes = T_STEPS * n_envs
    hps, tapes = [], []
    for _ in range(2):
        hp = C.c_void_p()
        D._lib.check(lib.dfrl_malloc_host(ctx.h, nbytes, C.byref(hp)))
        hps.append(hp)
        tapes.append(np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nbytes,)))
    rng = np.random.default_rng(rank)
    streams = [(rng.random(nbytes) < 0.4).astype(np.uint8) for _ in range(4)]

    def submit(i):
        tapes[i & 1][:] = streams[i % 4]         # the producer filling a pinned tape
        tr.rollout_raw(hps[i & 1], None, None)   # H2D of the tape + rollout
        tr.learn()
        tr.stats_begin()                         # D2H of the step's result (asynchronous)

    def run(n):
        s = None
        submit(0)
        for i in range(1, n):
            submit(i)
            s = tr.stats_end()                   # result of step i - 1 (tape (i - 1) & 1 is free again)
        return tr.stats_end() if n else s

    run(max(2, warmup))
    e2e_blocks = []
    for _ in range(repeats):
        barrier()
        t0 = time.perf_counter()
        s = run(steps)
        ctx.sync()
        e2e_blocks.append(max_over_ranks(time.perf_counter() - t0))
    dt = float(np.median(e2e_blocks))
    e2e_rate = n_envs * world * T_STEPS * steps / dt
    if sampler:
        sampler.__exit__()
    for hp in hps:
        D._lib.check(lib.dfrl_free_host(ctx.h, hp))
    res = {"value": value_rate, "ms_per_step": ms / steps, "launches_per_step": launches / steps,
           "repeats": repeats, "timed_ms_total": float(sum(blocks)), "block_ms_min_max": [float(min(blocks)), float(max(blocks))],
           "e2e": {"value": e2e_rate, "unit": "env-steps/s", "h2d_bytes_per_step": nbytes * world,
                   "d2h_bytes_per_step": 32 * world, "ms_per_step": 1e3 * dt / steps},
           "clocks": sampler.summary() if sampler else None, "stats": s,
           "objects": (tr, env, policy, value)}
    return res


def ctx_device(ctx):
    return int(os.environ.get("LOCAL_RANK", "0"))


def kernel_profile(D, ctx, tr, iters):
    """Per-kernel device time with CUDA events on the launching stream (separate pass: the event
    pairs add launch overhead, so this never overlaps the throughput measurement)."""
    lib = D._lib.lib
    tr.iterate(1)
    ctx.sync()
    D._lib.check(lib.dfrl_profile_enable(ctx.h, 1))
    tr.iterate(iters)
    buf = C.create_string_buffer(1 << 16)
    D._lib.check(lib.dfrl_profile_report(ctx.h, buf, len(buf)))
    D._lib.check(lib.dfrl_profile_enable(ctx.h, 0))
    prof = {}
    for line in buf.value.decode().strip().splitlines():
        name, n, ms = line.rsplit(" ", 2)
        prof[name] = {"launches": int(n) / iters, "ms": float(ms) / iters}
    return prof


def net_flops(dims):
    return sum(2 * a * b for a, b in zip(dims[:-1], dims[1:]))


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel, from the committed
    `ncu --set full` summary of the same bench command (profiles/ncu_summary.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if not os.path.exists(p):
        return None
    for k, v in json.load(open(p)).get("kernels", {}).items():
        if kernel_substr in k:
            return v.get("dram_bytes_read", 0) + v.get("dram_bytes_write", 0)
    return None


def roofline_from_profile(prof, n_envs, peaks):
    """Roofline of the dominant kernel and a per-kernel table (DESIGN.md section 'Kernels').

    Algorithmic work per launch counts only rows the algorithm needs: the N*T recorded start rows
    (forward + dX + dW = 3 F per row for a learner pass) and N end rows for the critic / GAE value
    evaluations; the reference additionally pushes the end rows through every policy pass with a
    zero gradient, which is not counted here."""
    rows = n_envs * T_STEPS
    fp, fv = net_flops(POLICY_DIMS), net_flops(VALUE_DIMS)
    P = 2 * 8 + 2
    work = {  # kernel substring -> (algorithmic FLOPs per launch, algorithmic HBM bytes per launch)
        "fused_policy_step": (rows * 3 * fp, rows * (P + 1 + 4 + 4 * 8)),
        "CRITIC_STEP": (rows * 3 * fv + n_envs * fv, rows * (P + 2 + 4) + n_envs * P),  # fused_critic_kernel
        "CRITIC_GAE": ((rows + n_envs) * fv, rows * (P + 2 + 4) + n_envs * P),          # fused_critic_kernel
        "fused_critic_step": (rows * 3 * fv + n_envs * fv, rows * (P + 2 + 4) + n_envs * P),
        "fused_gae": ((rows + n_envs) * fv, rows * (P + 2 + 4) + n_envs * P),
        "fused_rollout": (rows * fp, rows * (P + 2 + 4 * 8) + n_envs * (2 * P + 16)),
        "fused_reduce_partials": (0, None),
    }
    total = sum(v["ms"] for v in prof.values())
    table = {}
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        short = name.split("<")[0].strip("() ")
        for mode in ("CRITIC_STEP", "CRITIC_GAE"):  # two instantiations of one kernel template
            if mode in name:
                short += "<" + mode + ">"
        w = next((w for k, w in work.items() if k in name), None)
        per_launch_ms = v["ms"] / max(v["launches"], 1e-9)
        row = {"launches_per_step": v["launches"], "ms_per_step": round(v["ms"], 4),
               "us_per_launch": round(1e3 * per_launch_ms, 2), "share_of_step": round(v["ms"] / total, 4) if total else None}
        if w and w[0]:
            row["algorithmic_gflop_per_launch"] = round(w[0] / 1e9, 3)
            row["tflops"] = round(w[0] / (per_launch_ms * 1e-3) / 1e12, 2)
            row["frac_of_bf16_peak"] = round(row["tflops"] / peaks["bf16_tflops_sustained"], 4)
        if w and w[1]:
            row["algorithmic_mb_per_launch"] = round(w[1] / 1e6, 3)
            row["gbs"] = round(w[1] / (per_launch_ms * 1e-3) / 1e9, 1)
            row["frac_of_hbm_peak"] = round(row["gbs"] / peaks["hbm_gbs"], 4)
        table[short] = row
    top_name, top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    tw = next((w for k, w in work.items() if k in top_name), (0, None))
    per_launch_s = top["ms"] / max(top["launches"], 1e-9) * 1e-3
    achieved = tw[0] / per_launch_s / 1e12 if per_launch_s > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    return {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": ncu_traffic(top_name.split("<")[0].strip("() ")),
        "kernel": top_name.split("<")[0].strip("() "),
        "kernel_us_per_launch": 1e6 * per_launch_s, "kernel_launches_per_step": top["launches"],
        "kernel_ms_per_step": top["ms"], "all_kernels_ms_per_step": total,
        "share_of_step": top["ms"] / total if total else None,
        "algorithmic_flops_per_launch": tw[0], "algorithmic_bytes_per_launch": tw[1],
        "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['source']}): the kernel is timed inside a long step",
        "note": "FP32-grade results on the bf16 tensor pipe cost 3 tcgen05.mma per product (hi.hi + hi.lo + lo.hi): "
                "the pipe executes 3x the algorithmic FLOPs counted here",
        "kernels": table,
    }


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    import dependence_free_rl_b200 as D
    dist = None
    nccl_id = None
    if world > 1:
        # NCCL_DEBUG output (the driver counts ranks from it) goes to stderr: stdout carries ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("gloo", rank=rank, world_size=world)
        dist = dist_mod
        obj = [D.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]
    # (NCCL prints its version banner to fd 1 during ncclCommInitRank whatever NCCL_DEBUG_FILE says:
    #  fd 1 points at stderr while the communicator is created)
    sys.stdout.flush()
    saved_fd1 = os.dup(1) if world > 1 else -1
    if world > 1:
        os.dup2(2, 1)
    try:
        ctx = D.Context(local_rank, world, rank, nccl_id)
    finally:
        if world > 1:
            os.dup2(saved_fd1, 1)
            os.close(saved_fd1)
    exchange = "none (single rank)"
    if world > 1:
        exchange = "ncclAllReduce + optimizer kernel"
        if not args.no_p2p:
            # fused exchange: IPC handles of every rank's buffer -> peers mapped over NVLink
            handles = [None] * world
            dist.all_gather_object(handles, ctx.p2p_export())
            ctx.p2p_attach(handles)
            exchange = ("peer-memory PUSH over NVLink (flag-in-data 8-byte stores into every peer's buffer) fused with "
                        "the gradient reduction and the optimizer: one kernel per update, no NCCL call on the data path")
    peaks = load_peaks()

    main = measure(D, ctx, dist, args, args.envs_per_gpu, world, rank, args.steps, args.warmup, True)
    tr = main["objects"][0]
    prof = kernel_profile(D, ctx, tr, 3)
    roof = roofline_from_profile(prof, args.envs_per_gpu, peaks)
    for o in main["objects"]:
        o.close()

    extra = {}
    if world == 1 and not args.no_c5:
        # BASELINE configs[4] (GEMM-bound): 32 bins, 128-256-256-256-{32,1} nets; layered path whose
        # dense products run on the tcgen05 GEMMs of csrc/gemm_umma.cu
        n5 = args.c5_envs
        tr5, env5, p5, v5 = make_trainer(D, ctx, n5, 0, n5 * T_STEPS, pdims=C5_POLICY_DIMS, vdims=C5_VALUE_DIMS, n_bins=32)
        tr5.iterate(2)
        ctx.sync()
        ctx.timer_start()
        tr5.iterate(5)
        ms5 = ctx.timer_stop() / 5
        prof5 = kernel_profile(D, ctx, tr5, 2)
        f5 = flops_per_env_step(C5_POLICY_DIMS, C5_VALUE_DIMS) * n5 * T_STEPS
        gemm_ms = sum(v["ms"] for k, v in prof5.items() if "umma_gemm" in k)
        extra["c5_32bins_256hidden"] = {
            "value": n5 * T_STEPS / (ms5 * 1e-3), "unit": "env-steps/s", "envs": n5, "ms_per_step": ms5,
            "algorithmic_tflops_whole_step": f5 / (ms5 * 1e-3) / 1e12,
            "umma_gemm_ms_per_step": gemm_ms,
            "umma_gemm_tflops": f5 / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
            "umma_gemm_frac_of_bf16_peak": f5 / (gemm_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"] if gemm_ms else None,
            "per_kernel_ms": {k.split("(")[-1].split("<")[0].strip("() "): round(v["ms"], 3)
                              for k, v in sorted(prof5.items(), key=lambda kv: -kv[1]["ms"])[:8]}}
        for o in (tr5, env5, p5, v5):
            o.close()
    if world == 1 and not args.no_c2:
        # BASELINE configs[2]: online actor-critic (one critic step, GAE, one policy step with
        # A (p - onehot)), 65 536 envs, T = 8 (ac_training.cc:30), SHARED-TRUNK policy / value net
        # 32-64-64-{8, 1} (P = 6 857; dfrl_mlp_create_shared). The reference's sequential `model` cannot
        # express it and ac_training.cc builds two nets: that variant is reported beside it.
        n3, T3 = 65536, 8
        extra["c3_actor_critic_65536_envs"] = quick_rate(
            D, ctx, make_trainer(D, ctx, n3, 0, n3 * T3, algo=D.ACTOR_CRITIC, work=T3, last=D.SOFTMAX_CE, shared=True), n3, T3, 5, 50)
        extra["c3_actor_critic_65536_envs"]["net"] = "shared trunk 32-64-64, heads 64-8 (softmax-CE) and 64-1"
        extra["c3_actor_critic_separate_nets"] = quick_rate(
            D, ctx, make_trainer(D, ctx, n3, 0, n3 * T3, algo=D.ACTOR_CRITIC, work=T3, last=D.SOFTMAX_CE), n3, T3, 5, 50)
    if world == 1 and not args.no_extra:
        n = args.envs_per_gpu
        # T = 32 steps per iteration (SURVEY section 8d "also report T = 32"): 4 envs per learner tile
        extra["ppo_T32"] = quick_rate(D, ctx, make_trainer(D, ctx, n // 8, 0, (n // 8) * 32, work=32), n // 8, 32, 3, 20)
        # Adam (north_star names it; the reference apps use SGD): device-side step counter, same CUDA graph
        extra["ppo_adam"] = quick_rate(D, ctx, make_trainer(D, ctx, n, 0, n * T_STEPS, opt=D.ADAM, lr_scale=0.1), n, T_STEPS, 5, 50)
        # the reference's own default nets (ppo_training.cc:10-26: conv1d 4-128-64-1 softmax policy over the
        # 8 bins, FC 32-64-32-1 critic) on the fused tcgen05 kernels (fused_conv.cuh + the fused critic
        # kernels, CUDA-graphed learn phase), at 4096 envs (beside the reference's CPU rate for the same
        # nets below) and at the headline batch size
        # Two implementations of the conv1d policy: the tcgen05 kernels (every (sample, bin) row through the MLP) and,
        # from 8 192 envs x 4 steps on, the net evaluated on its FINITE input domain (conv_table.cuh: a conv1d_1 net
        # sees only (bin, item) pairs -- 6 561 distinct rows for 8 x 8 bins -- so the forward pass is a table lookup and
        # the backward pass one pass over the entries that occur, weighted by a fixed-point histogram of dY: the same
        # function and gradient, regrouped). Both are reported at the headline size; DFRL_CONV_TABLE picks one.
        for n_ref, key, table in ((4096, "reference_nets_4096_envs", None), (n, "reference_nets_%d_envs" % n, None),
                                  (n, "reference_nets_%d_envs_tensor_core_kernels" % n, "0")):
            if table is not None:
                os.environ["DFRL_CONV_TABLE"] = table
            row = quick_rate(
                D, ctx, make_trainer(D, ctx, n_ref, 0, n_ref * T_STEPS, player=D.conv_layers([4, 128, 64, 1], D.SOFTMAX),
                                     vlayer=D.fc_layers([32, 64, 32, 1])), n_ref, T_STEPS, 3, 20)
            os.environ.pop("DFRL_CONV_TABLE", None)
            on_table = table != "0" and n_ref * T_STEPS >= 32768
            row["policy_path"] = ("finite-domain table + fixed-point dY histogram (conv_table.cuh), HBM / L2-bound"
                                  if on_table else "tcgen05 kernels (fused_conv.cuh), FP16 hi / lo pairs")
            row["algorithmic_flop_per_env_step"] = 2296208   # SURVEY 8d, if every row went through the MLP
            if not on_table:
                row["algorithmic_tflops_whole_step"] = row["value"] * 2296208 / 1e12
                row["frac_of_bf16_peak"] = row["algorithmic_tflops_whole_step"] / peaks["bf16_tflops_sustained"]
            extra[key] = row
        # The two learners whose policy steps stay on the layered kernels: KL-PPO (kl_ppo_learner,
        # policy_gradient.h:310-335: end rows join the policy pass; critic step / GAE fused, beta adapted on the
        # device between the steps, learn phase replayed as a CUDA graph) and REINFORCE with the reference's FC policy
        # 32-256-128-8 + softmax-CE (pg_training.cc:12-18; whole episodes per iteration: env-steps
        # counted by the device statistics). 4096 envs each; wall clock around synchronous iterations.
        def layered_rate(objs, n_envs, warm, iters):
            tr = objs[0]
            tr.iterate(warm)
            s0 = tr.stats()["env_steps"]
            l0 = ctx.launches()
            t0 = time.perf_counter()
            tr.iterate(iters)
            s1 = tr.stats()["env_steps"]          # (synchronises)
            dt = time.perf_counter() - t0
            launches = (ctx.launches() - l0) / iters
            for o in (objs[0], objs[1], objs[3], objs[2]):
                if o is not None:
                    o.close()
            return {"value": (s1 - s0) / dt, "unit": "env-steps/s", "ms_per_step": 1e3 * dt / iters, "envs": n_envs,
                    "env_steps_per_iter": (s1 - s0) / iters, "launches_per_step": launches, "path": "layered kernels"}
        extra["kl_ppo_4096_envs"] = layered_rate(make_trainer(D, ctx, 4096, 0, 4096 * T_STEPS, algo=D.KL_PPO), 4096, 3, 20)
        pg = make_trainer(D, ctx, 4096, 0, 4096 * 13, algo=D.REINFORCE, work=1, pdims=[32, 256, 128, 8], last=D.SOFTMAX_CE)
        extra["reinforce_refnet_4096_envs"] = layered_rate(pg, 4096, 2, 10)
        extra["hbm_kernels"] = hbm_kernel_rows(D, ctx, peaks)
    if world == 1 and not args.no_c2:
        # BASELINE configs[1] verbatim: PPO, 4096 parallel envs, 1 GPU (latency-bound size)
        c2 = measure(D, ctx, None, args, 4096, 1, 0, max(args.steps, 50), max(args.warmup, 5), False)
        for o in c2["objects"]:
            o.close()
        extra["c2_4096_envs"] = {"value": c2["value"], "unit": "env-steps/s", "ms_per_step": c2["ms_per_step"],
                                 "e2e": c2["e2e"], "launches_per_step": c2["launches_per_step"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_extra:
        rows = cpu_reference_rows(os.cpu_count() or 1)
        if rows:
            extra["cpu_reference_rows"] = rows
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ref_envs = args.ref_envs if args.ref_envs > 0 else 256
        rate, secs, kind, used = cpu_reference_rate(ref_envs, args.ref_iters, cores)
        cpu = {"value": rate, "unit": "env-steps/s", "cores": used, "kind": kind,
               "sample": f"{args.ref_iters} PPO iterations of {ref_envs} envs x {T_STEPS} steps ({secs:.1f} s), "
                         f"same nets; rollouts on {used} threads, learner single-threaded as in the reference"}

    if rank == 0:
        clocks = main["clocks"] or {}
        line = {
            "metric": "ppo_binpacking_env_steps_per_sec", "value": main["value"], "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 via bf16x3 (fp32 operands split hi + lo, 3 tcgen05 bf16 products per fp32 product, fp32 accumulation in TMEM; ~2^-17 relative)",
            "repeats": main["repeats"], "timed_ms_total": main["timed_ms_total"], "block_ms_min_max": main["block_ms_min_max"],
            "data": "synthetic", "config": {**workload_config(args), "gradient_exchange": exchange},
            "e2e": main["e2e"], "gpu_launches": int(round(main["launches_per_step"] * args.steps)),
            "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"),
                       "reasons": clocks.get("reasons", []), "samples": clocks.get("samples", 0)},
            "roofline": roof, "cpu_baseline": cpu,
            "train_stats": main["stats"], **extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)   # ~0.4 s timed at ~1 ms per PPO iteration
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=131072)
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--ref-envs", type=int, default=0,
                    help="bounded CPU sample: envs per PPO iteration (0: sized from --steps for ~1 min of CPU work)")
    ap.add_argument("--ref-iters", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--c5-envs", type=int, default=131072)  # C5 at 131 072 envs / GPU
    ap.add_argument("--min-timed-s", type=float, default=0.6,
                    help="repeat the --steps block until this much device time has been measured; report the median block")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary rows (T=32, Adam, reference nets, HBM kernels, C1)")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL all-reduce instead of the fused peer-memory exchange")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
