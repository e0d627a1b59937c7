"""Shared-trunk policy / value net (BASELINE configs[2]: "online actor-critic, 65536 envs,
shared-trunk policy/value MLP"). The reference's `model` is strictly sequential (nn.h:467-542) and
its mains build two nets (ac_training.cc:9-25), so there is no reference trace for this extension:
the oracle (dfrl_oracle.c, explicit parameter offsets) is hand-composed from the reference's layer
forward / backward / gradient and is checked here against an INDEPENDENT composition written with
the sequential-net primitives only (CPU); the GPU path is then checked against the oracle.

Semantics (what a reference user gets when both `model`s of actor_critic_learner hold the same trunk
layers): the learner's own order -- critic step, advantages with the UPDATED net, actor step
(policy_gradient.h:159-185) -- each step moving the trunk through its own head's loss."""
import numpy as np
import pytest

import flipcheck
from refcases import close

TRUNK, NB = [32, 64, 64], 8
P_TRUNK = 32 * 64 + 64 + 64 * 64 + 64          # 6272
P_PI, P_V = 64 * 8 + 8, 64 + 1                 # 520, 65
P_ALL = P_TRUNK + P_PI + P_V                   # 6857 (C3, P = 6 857)


def shared_nets(orc, last):
    """(policy net, value net) over ONE vector [trunk | pi head | v head]."""
    pl = [(orc.DENSE, 32, 64), (orc.RELU, 0, 0), (orc.DENSE, 64, 64), (orc.RELU, 0, 0), (orc.DENSE, 64, 8), (last, 0, 0)]
    vl = [(orc.DENSE, 32, 64), (orc.RELU, 0, 0), (orc.DENSE, 64, 64), (orc.RELU, 0, 0), (orc.DENSE, 64, 1)]
    o1, o2 = 0, 32 * 64 + 64
    pnet = orc.Net(pl, 32).with_offsets([o1, None, o2, None, P_TRUNK, None], P_ALL)
    vnet = orc.Net(vl, 32).with_offsets([o1, None, o2, None, P_TRUNK + P_PI], P_ALL)
    return pnet, vnet


def realistic(rng):
    p = np.zeros(P_ALL, np.float32)
    for off, a, b in ((0, 32, 64), (32 * 64 + 64, 64, 64), (P_TRUNK, 64, 8), (P_TRUNK + P_PI, 64, 1)):
        p[off:off + a * b] = (rng.standard_normal(a * b) * 0.01).astype(np.float32)   # nn.h:12-14
    return p


def safe(rng):
    """Hidden biases of +-5: no relu pre-activation near zero (strict 1e-4 comparison possible)."""
    p = np.zeros(P_ALL, np.float32)
    for off, a, b, hidden in ((0, 32, 64, True), (32 * 64 + 64, 64, 64, True), (P_TRUNK, 64, 8, False), (P_TRUNK + P_PI, 64, 1, False)):
        p[off:off + a * b] = (rng.standard_normal(a * b) * 0.05).astype(np.float32)
        bias = (rng.standard_normal(b) * 0.05).astype(np.float32)
        if hidden:
            bias += np.where(np.arange(b) % 2 == 0, 5.0, -5.0).astype(np.float32)
        p[off + a * b:off + a * b + b] = bias
    return p


def test_param_count():
    assert P_ALL == 6857 == 2112 + 4160 + 520 + 65


def test_oracle_shared_trunk_against_independent_composition():
    from oracle import orc
    orc.build()
    n, T, gamma, lam = 96, 8, 0.99, 0.95
    rng = np.random.default_rng(3)
    pnet, vnet = shared_nets(orc, orc.SOFTMAX_CE)
    p0 = safe(rng)
    ecfg = orc.env_cfg(NB)
    st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
    items = rng.integers(0, 2, (T, n)).astype(np.uint8)
    st_before = st.copy()
    ro = orc.rollout(ecfg, st, pnet, p0, T, 0, items, u=rng.random((T, n)))
    plr, vlr = 1e-7, 1e-7
    L = orc.Learner(orc.train_cfg(orc.ACTOR_CRITIC, T, policy_lr=plr, value_lr=vlr), ecfg, pnet, p0, vnet, None)
    out = L.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
    assert L.vparams is L.pparams

    # ---- independent composition with sequential nets only
    seq_p, seq_v = orc.fc_net(TRUNK + [8], orc.SOFTMAX_CE), orc.fc_net(TRUNK + [1])
    pi = lambda p: np.concatenate([p[:P_TRUNK], p[P_TRUNK:P_TRUNK + P_PI]])
    vv = lambda p: np.concatenate([p[:P_TRUNK], p[P_TRUNK + P_PI:]])
    done, act = ro["done"].astype(bool), ro["action"]
    obs_s = orc.obs_encode(ro["state"].transpose(1, 0, 2).reshape(18, T * n), NB)
    # end states: terminal (bin[a] -= item, item kept) when done, live state at the last step
    end = ro["state"].copy()
    for t in range(T):
        for i in range(n):
            if done[t, i]:
                a = act[t, i]
                end[t, 2 * a, i] -= end[t, 16, i]
                end[t, 2 * a + 1, i] -= end[t, 17, i]
            elif t == T - 1:
                end[t, :, i] = st[:, i]
    obs_e = orc.obs_encode(end.transpose(1, 0, 2).reshape(18, T * n), NB)

    def values(p):
        vs = orc.net_eval(seq_v, vv(p), obs_s).reshape(T, n)
        ve = orc.net_eval(seq_v, vv(p), obs_e).reshape(T, n)
        vn = np.where(done | (np.arange(T)[:, None] == T - 1), ve, np.roll(vs, -1, axis=0))
        return vs, vn
    vs, vn = values(p0)
    target = (1.0 - done) + gamma * vn                                # not masked at terminals (quirk 6)
    gv, _ = orc.net_forward_gradient(seq_v, vv(p0), obs_s, (vs - target).reshape(-1, 1).astype(np.float32))
    p1 = p0.copy()
    p1[:P_TRUNK] -= vlr * gv[:P_TRUNK]
    p1[P_TRUNK + P_PI:] -= vlr * gv[P_TRUNK:]
    vs, vn = values(p1)                                               # UPDATED net (quirk 7)
    vn_adv = np.where(done, 0.0, vn)
    delta = (1.0 - done) + gamma * vn_adv - vs
    adv = np.zeros((T, n), np.float32)
    for t in range(T - 1, -1, -1):
        nxt = 0.0 if t == T - 1 else np.where(done[t], 0.0, adv[t + 1])
        adv[t] = delta[t] + gamma * lam * nxt
    probs = orc.net_eval(seq_p, pi(p1), obs_s)
    dy = orc.loss_grad(orc.LOSS_SOFTMAX_LOG, probs, act.reshape(-1), adv.reshape(-1))
    gp, _ = orc.net_forward_gradient(seq_p, pi(p1), obs_s, dy)
    p2 = p1.copy()
    p2[:P_TRUNK + P_PI] -= plr * gp
    close(out["adv"], adv, what="advantages")
    close(out["value_grad"][:P_TRUNK], gv[:P_TRUNK], what="critic gradient, trunk")
    close(out["value_grad"][P_TRUNK + P_PI:], gv[P_TRUNK:], what="critic gradient, value head")
    assert np.all(out["value_grad"][P_TRUNK:P_TRUNK + P_PI] == 0)
    close(out["policy_grads"][0][:P_TRUNK + P_PI], gp, what="actor gradient")
    assert np.all(out["policy_grads"][0][P_TRUNK + P_PI:] == 0)
    close(L.pparams, p2, what="parameters after one iteration")
    assert np.any(p2[:P_TRUNK] != p1[:P_TRUNK]) and np.any(p1[:P_TRUNK] != p0[:P_TRUNK])  # the trunk moved twice


def _gpu_models(D, ctx, last):
    policy = D.Model(ctx, D.fc_layers(TRUNK + [8], last), 32)
    value = D.Model.shared(policy, 4, [(D.DENSE, 64, 1)])
    assert policy.n_params == value.n_params == P_ALL
    return policy, value


@pytest.mark.gpu
@pytest.mark.parametrize("algo_name,n,T,ctas", [("ac", 1100, 8, 2), ("ppo", 700, 4, 3), ("ac", 300, 8, 0)])
def test_shared_trunk_trainer_vs_oracle(D, ctx, orc, algo_name, n, T, ctas):
    algo, oalgo, last, olast = ((D.ACTOR_CRITIC, orc.ACTOR_CRITIC, D.SOFTMAX_CE, orc.SOFTMAX_CE) if algo_name == "ac"
                                else (D.PPO, orc.PPO, D.SOFTMAX, orc.SOFTMAX))
    rng = np.random.default_rng(5)
    pnet, vnet = shared_nets(orc, olast)
    p0 = safe(rng)
    policy, value = _gpu_models(D, ctx, last)
    policy.set_parameters(p0)
    assert np.array_equal(value.parameters(), p0)
    ecfg = orc.env_cfg(NB)
    st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
    env = D.Environment(ctx, n)
    env.set_state(st)
    tr = D.Trainer(ctx, env, policy, value, algo=algo, work=T, policy_lr=2e-8, value_lr=2e-8)
    if ctas:
        D._lib.check(D._lib.lib.dfrl_debug_set_fused_ctas(tr.h, ctas))
    L = orc.Learner(orc.train_cfg(oalgo, T, policy_lr=2e-8, value_lr=2e-8), ecfg, pnet, p0, vnet, None)
    for it in range(4):  # (the learn phase replays as a CUDA graph from the third iteration on)
        items = rng.integers(0, 2, (T, n)).astype(np.uint8)
        u = rng.random((T, n))
        ro = orc.rollout(ecfg, st, pnet, L.pparams, T, 0, items, u=u)
        tr.rollout(items=items, u=u)
        assert np.array_equal(tr.read(D.F_REC_ACTION), ro["action"])
        assert np.array_equal(tr.read(D.F_REC_DONE), ro["done"])
        assert np.array_equal(env.state(), st)
        out = L.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
        tr.learn()
        close(tr.read(D.F_ADVANTAGE), out["adv"], what="adv")
        vg = tr.read(D.F_VALUE_GRAD)
        close(vg, out["value_grad"], what="value gradient")
        assert np.all(vg[P_TRUNK:P_TRUNK + P_PI] == 0)
        pg = tr.read(D.F_POLICY_GRAD_LOG)
        close(pg, out["policy_grads"], what="policy gradients")
        assert np.all(pg[:, P_TRUNK + P_PI:] == 0)
        close(policy.parameters(), L.pparams, what="family parameters")
        assert np.array_equal(policy.parameters(), value.parameters())
    tr.close(); env.close(); value.close(); policy.close()


@pytest.mark.gpu
def test_shared_trunk_layered_path_and_family_rules(D, ctx, orc):
    """fused=0: the layered kernels (any net shape) with a shared trunk; plus the family rules of
    dfrl_mlp_create_shared (init of a sharer touches its head only, destroy order)."""
    n, T = 200, 4
    rng = np.random.default_rng(8)
    pnet, vnet = shared_nets(orc, orc.SOFTMAX)
    policy, value = _gpu_models(D, ctx, D.SOFTMAX)
    policy.init_parameters(1)
    before = policy.parameters()
    value.init_parameters(2)
    after = policy.parameters()
    assert np.array_equal(before[:P_TRUNK + P_PI], after[:P_TRUNK + P_PI]) and np.any(after[P_TRUNK + P_PI:-1] != 0)
    with pytest.raises(D._lib.DfrlError, match="share"):
        policy.close()
    p0 = safe(rng)
    value.set_parameters(p0)
    ecfg = orc.env_cfg(NB)
    st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
    env = D.Environment(ctx, n)
    env.set_state(st)
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=2e-8, value_lr=2e-8, fused=0)
    L = orc.Learner(orc.train_cfg(orc.PPO, T, policy_lr=2e-8, value_lr=2e-8), ecfg, pnet, p0, vnet, None)
    for it in range(2):
        items = rng.integers(0, 2, (T, n)).astype(np.uint8)
        u = rng.random((T, n))
        ro = orc.rollout(ecfg, st, pnet, L.pparams, T, 0, items, u=u)
        tr.rollout(items=items, u=u)
        assert np.array_equal(tr.read(D.F_REC_ACTION), ro["action"])
        out = L.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
        tr.learn()
        close(tr.read(D.F_VALUE_GRAD), out["value_grad"], what="value gradient")
        close(tr.read(D.F_POLICY_GRAD_LOG), out["policy_grads"], what="policy gradients")
        close(policy.parameters(), L.pparams, what="family parameters")
    tr.close(); env.close(); value.close(); policy.close()
