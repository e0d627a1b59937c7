# Debug aid: conv1d policy step vs the oracle, per parameter block, for a few batch sizes / CTA caps.
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import dependence_free_rl_b200 as D
from oracle import orc
from test_gpu_parity import _safe_params, _conv_safe_params
ctx = D.Context(0)
def run(n, T, cap, dims=[4, 128, 64, 1], vdims=[32, 64, 32, 1], algo="ppo"):
    B = 8
    rng = np.random.default_rng(5)
    last = D.SOFTMAX if algo == "ppo" else D.SOFTMAX_CE
    pl, vl = D.conv_layers(dims, last), D.fc_layers(vdims)
    pnet, vnet = orc.Net(pl, 32), orc.Net(vl, 32)
    pp, vp = _conv_safe_params(dims, 11), _safe_params(vdims, 12)
    policy, value = D.Model(ctx, pl, 32), D.Model(ctx, vl, 32)
    policy.set_parameters(pp); value.set_parameters(vp)
    ecfg = orc.env_cfg(B)
    st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
    env = D.Environment(ctx, n, n_bins=B); env.set_state(st)
    A, OA = (D.PPO, orc.PPO) if algo == "ppo" else (D.ACTOR_CRITIC, orc.ACTOR_CRITIC)
    tr = D.Trainer(ctx, env, policy, value, algo=A, work=T, policy_lr=2e-8, value_lr=2e-8, action_mode=D.ACT_SAMPLE)
    if cap:
        D._lib.check(D._lib.lib.dfrl_debug_set_fused_ctas(tr.h, cap))
    lr = orc.Learner(orc.train_cfg(OA, T, policy_lr=2e-8, value_lr=2e-8), ecfg, pnet, pp, vnet, vp)
    items = rng.integers(0, 2, (T, n)).astype(np.uint8); u = rng.random((T, n))
    ro = orc.rollout(ecfg, st, pnet, lr.pparams, T, 0, items, u=u)
    tr.rollout(items=items, u=u)
    print(f"n={n} T={T} cap={cap}: actions equal {np.array_equal(tr.read(D.F_REC_ACTION), ro['action'])}, "
          f"probs err {np.max(np.abs(tr.read(D.F_REC_PROBS) - ro['probs'])):.2e}")
    out = lr.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
    tr.learn()
    g, w = tr.read(D.F_POLICY_GRAD_LOG), out["policy_grads"]
    d1, d2 = dims[1], dims[2]
    blocks = [("W1", 0, 4 * d1), ("b1", 4 * d1, 5 * d1), ("W2", 5 * d1, 5 * d1 + d1 * d2), ("b2", 5 * d1 + d1 * d2, 5 * d1 + d1 * d2 + d2),
              ("W3", 5 * d1 + d1 * d2 + d2, 5 * d1 + d1 * d2 + 2 * d2), ("b3", 5 * d1 + d1 * d2 + 2 * d2, 5 * d1 + d1 * d2 + 2 * d2 + 1)]
    for e in range(g.shape[0]):
        print("  epoch", e, " ".join(f"{nm} {np.linalg.norm(g[e, a:b] - w[e, a:b]) / (np.linalg.norm(w[e, a:b]) + 1e-30):.1e}" for nm, a, b in blocks))
        if e == 0:
            print("     norms gpu/oracle:", " ".join(f"{nm} {np.linalg.norm(g[e, a:b]):.2e}/{np.linalg.norm(w[e, a:b]):.2e}" for nm, a, b in blocks))
    for o in (tr, env, policy, value): o.close()
for n, cap in [(4, 1), (64, 1), (300, 2)]:
    run(n, 4, cap)
run(300, 4, 2, algo="ac")
run(300, 4, 2, dims=[4, 64, 32, 1], algo="ppo")
run(200, 8, 2, dims=[4, 64, 32, 1], algo="ac")

def run_golden(name="ppo_refnet_conv"):
    import refcases
    c = refcases.load_case(name)
    algo, n, work = int(c["algo"]), int(c["n_envs"]), int(c["work"])
    for fused in (1, 0):
        policy = D.Model(ctx, c["policy_layers"], 32); policy.set_parameters(c["pparams0"])
        value = D.Model(ctx, c["value_layers"], 32); value.set_parameters(c["vparams0"])
        env = D.Environment(ctx, n); env.set_state(refcases.initial_state(c["steps"], n))
        tr = D.Trainer(ctx, env, policy, value, algo=algo, work=work, policy_lr=float(c["plr"]), value_lr=float(c["vlr"]),
                       action_mode=D.ACT_FORCED, fused=fused)
        rec = refcases.records_from_steps(c["steps"], 0, n)
        tr.rollout(items=rec["items"], actions=rec["action"])
        print(f"golden {name} fused={fused} coverage={tr.fused_coverage()}: p_old max err {np.max(np.abs(tr.read(D.F_REC_PROBS) - rec['p_old'])):.2e}")
        tr.learn()
        g = tr.read(D.F_POLICY_GRAD_LOG)
        d1, d2 = 128, 64
        blocks = [("W1", 0, 4 * d1), ("b1", 4 * d1, 5 * d1), ("W2", 5 * d1, 5 * d1 + d1 * d2), ("b2", 5 * d1 + d1 * d2, 5 * d1 + d1 * d2 + d2),
                  ("W3", 5 * d1 + d1 * d2 + d2, 5 * d1 + d1 * d2 + 2 * d2), ("b3", 5 * d1 + d1 * d2 + 2 * d2, 5 * d1 + d1 * d2 + 2 * d2 + 1)]
        for e in range(4):
            w = c["policy_grads"][e]
            print("  epoch", e, f"all {np.linalg.norm(g[e] - w) / np.linalg.norm(w):.1e}", " ".join(f"{nm} {np.linalg.norm(g[e, a:b] - w[a:b]) / (np.linalg.norm(w[a:b]) + 1e-30):.1e}" for nm, a, b in blocks))
        adv = tr.read(D.F_ADVANTAGE)
        print("  adv", adv.reshape(-1)[:8])
        for o in (tr, env, policy, value): o.close()
run_golden()
