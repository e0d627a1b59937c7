"""Learning curve on the GPU path: PPO with the reference's own nets and rates (ppo_training.cc:9-31:
conv1d 4-128-64-1 softmax policy, FC 32-64-32-1 critic, SGD 1e-4 / 1e-5 scaled by 32 / rows because
gradients are SUMS over rows, T = 4, k = 4) at 4096 parallel envs, and with the fused C2 nets
(32-64-64-{8,1}); evaluated like the reference's main (argmax policy, ppo_training.cc:67-81) on 2048
fresh episodes. Beside it: the unmodified reference's own curve at equal round counts
(profiles/r02_convergence_ref_cpu.csv, tools/ref_convergence.py: 8 envs per round as ppo_training.cc).
    python tools/convergence.py [iterations] [every]   -> profiles/r02_convergence_gpu.csv"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dependence_free_rl_b200 as D  # noqa: E402


def curve(ctx, nets, n=4096, T=4, iters=5000, every=250, seed=1234, lr_scale=1.0):
    if nets == "reference":
        policy = D.Model(ctx, D.conv_layers([4, 128, 64, 1], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 32, 1]), 32)
    else:
        policy = D.Model(ctx, D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 64, 1]), 32)
    policy.init_parameters(seed)
    value.init_parameters(seed + 1)
    if nets != "reference":
        # He-scaled dense weights: the reference's N(0, 0.01) dense init (nn.h:12-14) leaves a 3-layer FC
        # net without signal for ~10^5 rounds (its own PPO main uses the He-initialised conv1d net)
        import numpy as np

        def he_fc(dims, s):
            r = np.random.default_rng(s)
            return np.concatenate([np.concatenate([(r.standard_normal(a * b) * np.sqrt(2.0 / a)).astype(np.float32),
                                                   np.zeros(b, np.float32)]) for a, b in zip(dims[:-1], dims[1:])])
        policy.set_parameters(he_fc([32, 64, 64, 8], seed))
        value.set_parameters(he_fc([32, 64, 64, 1], seed + 1))
    env = D.Environment(ctx, n, seed=seed)
    rows = n * T
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=lr_scale * 1e-4 * 32 / rows,
                   value_lr=lr_scale * 1e-5 * 32 / rows)
    out = []

    def evaluate(it):
        e = D.Environment(ctx, 1024, seed=99)
        mean, _ = D.eval_argmax(ctx, e, policy, 2)
        e.close()
        out.append((it, it * rows, mean))
    evaluate(0)
    for it in range(every, iters + 1, every):
        tr.iterate(every)
        evaluate(it)
    for o in (tr, env, value, policy):
        o.close()
    return out


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    every = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    ctx = D.Context(0)
    with open(os.path.join(ROOT, "profiles", "r02_convergence_gpu.csv"), "w") as f:
        f.write("nets,iterations,env_steps,mean_reward_argmax_2048ep\n")
        for nets in ("reference", "c2_fused_he_init"):
            for it, steps, mean in curve(ctx, nets, iters=iters, every=every):
                f.write(f"{nets},{it},{steps},{mean:.4f}\n")
                f.flush()
    print(open(os.path.join(ROOT, "profiles", "r02_convergence_gpu.csv")).read())
