"""GPU parity at the BASELINE sizes (-m gpu): the fused tcgen05 path through the C ABI against the
double-accumulating oracle (oracle/liboracle64mt.so = dfrl_oracle.c, -DORC_ACC_DOUBLE, OpenMP over
rows / outputs, bit-identical to the single-threaded build) with the REFERENCE'S initialisation
(dfrl_mlp_init_params = N(0, 0.01) weights, zero biases, nn.h:12-14) -- no 'safe' parameters.

  configs[1]  PPO, 4 096 envs x 4 steps, 3 iterations
  configs[3]  PPO, 131 072 envs x 4 steps (the per-GPU shard of 1 M envs), 1 iteration
  configs[2]  online actor-critic, 65 536 envs x 8 steps, 1 iteration -- with separate nets (the
              reference's ac_training.cc) and with the shared-trunk policy / value net

Bars: rollout records, sampled actions, done flags, env state bit-exact; probabilities,
advantages, targets, value gradient, parameters within 1e-4 (refcases.close, elementwise against
the vector's magnitude + norm-wise); policy gradients within 1e-4 after removing the
relu-ambiguous directions (tests/flipcheck.py explains why a noise-like gradient over 5 x 10^5
rows needs that, and why it cannot hide a wrong kernel)."""
import numpy as np
import pytest

import flipcheck
from refcases import close

pytestmark = pytest.mark.gpu

PD, VD = [32, 64, 64, 8], [32, 64, 64, 1]


def _setup(D, ctx, orc, n, T, algo, seed, conv=None, vd=None):
    """conv = [4, D1, D2, 1]: the reference's conv1d_1 policy (He init, nn.h:16-18) instead of the dense PD net."""
    B = 8
    shared = algo == "ac_shared"
    last = D.SOFTMAX if algo == "ppo" else D.SOFTMAX_CE
    pl, vl = (D.conv_layers(conv, last) if conv else D.fc_layers(PD, last)), D.fc_layers(vd or VD)
    policy = D.Model(ctx, pl, 32)
    policy.init_parameters(seed)
    if shared:  # one flat vector [trunk | policy head | value head] (tests/test_shared_trunk.py)
        import test_shared_trunk as sh
        value = D.Model.shared(policy, 4, [(D.DENSE, 64, 1)])
        value.init_parameters(seed + 1)
        pnet, vnet = sh.shared_nets(orc, orc.SOFTMAX_CE)
        pp, vp = policy.parameters(), None
    else:
        value = D.Model(ctx, vl, 32)
        value.init_parameters(seed + 1)
        pnet, vnet = orc.Net(pl, 32), orc.Net(vl, 32)
        pp, vp = policy.parameters(), value.parameters()
    if conv:
        assert abs(pp[:4 * conv[1]].std() - np.sqrt(2 / 4)) < 0.1 and np.all(pp[4 * conv[1]:5 * conv[1]] == 0)  # He init
    else:
        assert abs(pp[:32 * 64].std() - 0.01) < 1e-3 and np.all(pp[32 * 64:32 * 64 + 64] == 0)  # reference init
    rng = np.random.default_rng(seed + 2)
    ecfg = orc.env_cfg(B)
    st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
    for _ in range(6):  # spread the environments over their episodes
        orc.env_step(ecfg, st, rng.integers(0, B, n).astype(np.uint8), rng.integers(0, 2, n).astype(np.uint8))
    env = D.Environment(ctx, n)
    env.set_state(st)
    plr, vlr = 1e-4 * 32 / (n * T), 1e-5 * 32 / (n * T)  # reference rates for SUM gradients over 32 rows
    dalgo, oalgo = (D.PPO, orc.PPO) if algo == "ppo" else (D.ACTOR_CRITIC, orc.ACTOR_CRITIC)
    tr = D.Trainer(ctx, env, policy, value, algo=dalgo, work=T, policy_lr=plr, value_lr=vlr)
    lr = orc.Learner(orc.train_cfg(oalgo, T, policy_lr=plr, value_lr=vlr), ecfg, pnet, pp, vnet, vp, f64="mt")
    return dict(policy=policy, value=value, env=env, tr=tr, lr=lr, st=st, ecfg=ecfg, pnet=pnet, rng=rng, plr=plr)


def _iteration(D, orc, S, n, T, algo, it, conv=None):
    tr, lr, st, rng = S["tr"], S["lr"], S["st"], S["rng"]
    items = rng.integers(0, 2, (T, n)).astype(np.uint8)
    st0 = st.copy()
    pp0 = lr.pparams.copy()

    def oracle_rollout(u):
        s = st0.copy()
        ro = orc.rollout(S["ecfg"], s, S["pnet"], pp0, T, 0, items, u=u)
        ro["final"] = s
        return ro
    ro, u = flipcheck.nudge_uniforms(oracle_rollout, rng.random((T, n)), lambda r: r["probs"], rng)
    st[:] = ro["final"]
    tr.rollout(items=items, u=u)
    # ---- transitions: bit-exact
    assert np.array_equal(tr.read(D.F_REC_ACTION), ro["action"]), f"it {it}: sampled actions differ"
    assert np.array_equal(tr.read(D.F_REC_DONE), ro["done"]), f"it {it}: done flags differ"
    assert np.array_equal(tr.read(D.F_REC_STATE), ro["state"]), f"it {it}: recorded states differ"
    assert np.array_equal(S["env"].state(), st), f"it {it}: live env state differs"
    close(tr.read(D.F_REC_PROBS), ro["probs"], what=f"it {it} p_old")
    # ---- learn
    out = lr.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
    tr.learn()
    close(tr.read(D.F_ADVANTAGE), out["adv"], what=f"it {it} advantages")
    close(tr.read(D.F_VALUE_TARGET), out["targets"], what=f"it {it} targets")
    close(tr.read(D.F_VALUE_GRAD), out["value_grad"], what=f"it {it} value gradient")
    # ---- policy gradients, epoch by epoch, flip-aware
    obs = orc.obs_encode(ro["state"].transpose(1, 0, 2).reshape(18, T * n), 8)   # row k = t * n + i
    actions = ro["action"].reshape(-1).astype(np.int64)
    adv = out["adv"].reshape(-1).astype(np.float64)
    p_old_a = ro["probs"].reshape(-1, 8)[np.arange(T * n), actions].astype(np.float64)
    log = tr.read(D.F_POLICY_GRAD_LOG)
    kind = flipcheck.PPO if algo == "ppo" else flipcheck.AC
    if conv:
        return _conv_policy_reports(S, lr, out, log, obs, actions, adv, p_old_a, kind, conv, pp0, it)
    NP = 6792  # the policy net's own parameters lead the vector ([trunk | policy head | value head] when shared)
    # (shared trunk: the critic step has already moved the trunk when the actor step runs)
    params = pp0.copy() if algo != "ac_shared" else (lr.pparams + out["policy_grads"][0] * np.float32(S["plr"])).astype(np.float32)
    reports = []
    for e in range(log.shape[0]):
        def dout(rows, o):
            return flipcheck.policy_dlogits(o, actions[rows], adv[rows], p_old_a[rows], kind)
        Dm, cand = flipcheck.ambiguous_directions(obs, params[:NP], PD, dout)
        reports.append(flipcheck.flip_close(log[e][:NP], out["policy_grads"][e][:NP], Dm, what=f"it {it} policy gradient {e}"))
        assert np.all(log[e][NP:] == 0)
        params = (params - out["policy_grads"][e] * np.float32(S["plr"])).astype(np.float32)  # sgd, nn.h:622-625
    close(S["policy"].parameters(), lr.pparams, what=f"it {it} policy params")
    close(S["value"].parameters(), lr.vparams, what=f"it {it} value params")
    return reports


def _conv_policy_reports(S, lr, out, log, obs, actions, adv, p_old_a, kind, dims, pp0, it):
    """Flip-aware comparison for a conv1d_1 policy: the net is the dense net `dims` = [4, D1, D2, 1]
    over the (sample, bin) rows (nn.h:127-147), row = 8 * sample + bin; the gradient at a row's scalar
    output is its sample's logit gradient (float64 forward of the whole batch)."""
    R = obs.shape[0]
    xr = np.asarray(obs, np.float64).reshape(R * 8, 4)
    params = pp0.copy()
    reports = []
    for e in range(log.shape[0]):
        (W1, b1), (W2, b2), (W3, b3) = flipcheck.split_params(params, dims)
        h2 = np.maximum(np.maximum(xr @ W1.T + b1, 0.0) @ W2.T + b2, 0.0)
        logits = (h2 @ W3.T + b3).reshape(R, 8)
        DL = flipcheck.policy_dlogits(logits, actions, adv, p_old_a, kind).reshape(R * 8, 1)
        Dm, cand = flipcheck.ambiguous_directions(xr.astype(np.float32), params, dims, lambda rows, o: DL[rows])
        reports.append(flipcheck.flip_close(log[e], out["policy_grads"][e], Dm, what=f"it {it} conv policy gradient {e}"))
        params = (params - out["policy_grads"][e] * np.float32(S["plr"])).astype(np.float32)
    close(S["policy"].parameters(), lr.pparams, what=f"it {it} policy params")
    close(S["value"].parameters(), lr.vparams, what=f"it {it} value params")
    return reports


@pytest.mark.parametrize("algo,n,T,iters,conv,cap", [
    ("ppo", 4096, 4, 2, [4, 128, 64, 1], 0),   # ppo_training.cc:10-26 nets at the bench's reference_nets_4096_envs size
    ("ppo", 700, 4, 2, [4, 128, 64, 1], 3),    # 3 CTAs: 59 tiles per CTA (steady state of the tile loop, both observation slots)
    ("ppo", 37, 4, 1, [4, 128, 64, 1], 0),     # ragged: 148 samples = 9.25 tiles
    ("ac", 2048, 8, 2, [4, 64, 32, 1], 0),     # ac_training.cc:9-25 nets
    ("ppo", 500, 4, 1, [4, 64, 32, 1], 2),
])
@pytest.mark.parametrize("table", [0, 1])
def test_reference_nets_fused_vs_oracle64(D, ctx, orc, monkeypatch, algo, n, T, iters, conv, cap, table):
    """The reference's OWN default nets (conv1d_1 policy over the 8 bins, He init; critic 32-64-32-1,
    N(0, 0.01) init) against the double-accumulating oracle: transitions bit-exact, values / advantages /
    gradients within 1e-4. table = 0: the fused tcgen05 kernels (fused_conv.cuh + the fused critic kernels);
    table = 1: the policy on its finite input domain (conv_table.cuh: logit table, fixed-point histogram of dY,
    one backward pass), the default from 8 192 envs x 4 steps on."""
    monkeypatch.setenv("DFRL_CONV_TABLE", str(table))
    S = _setup(D, ctx, orc, n, T, algo, seed=11, conv=conv, vd=[32, 64, 32, 1])
    assert S["tr"].fused_covers_iteration()
    if cap:
        D._lib.check(D._lib.lib.dfrl_debug_set_fused_ctas(S["tr"].h, cap))
    p0 = S["policy"].parameters().copy()
    for it in range(iters):
        for e, r in enumerate(_iteration(D, orc, S, n, T, algo, it, conv=conv)):
            print(f"[{algo} conv {n}x{T}] iteration {it} policy gradient {e}: raw norm-wise error {r['raw']:.2e}, "
                  f"{r['ambiguous']} relu-ambiguous units ({r['flipped']} flipped), residual {r['residual']:.2e}")
            assert r["ambiguous"] < 0.35 * p0.size
    assert np.any(S["policy"].parameters() != p0)
    for k in ("tr", "env", "value", "policy"):
        S[k].close()


@pytest.mark.parametrize("algo,n,T,iters", [
    ("ppo", 4096, 4, 3),      # BASELINE configs[1]
    ("ppo", 131072, 4, 1),    # BASELINE configs[3]: one GPU's shard of the 1 M-env run
    ("ac", 65536, 8, 1),      # BASELINE configs[2]: online actor-critic, separate nets
    ("ac_shared", 65536, 8, 1),  # BASELINE configs[2] as worded: shared-trunk policy / value net
])
def test_fused_path_vs_oracle64_at_baseline_sizes(D, ctx, orc, algo, n, T, iters):
    S = _setup(D, ctx, orc, n, T, algo, seed=7)
    p0 = S["policy"].parameters().copy()
    total_amb = 0
    for it in range(iters):
        for e, r in enumerate(_iteration(D, orc, S, n, T, algo, it)):
            print(f"[{algo} {n}x{T}] iteration {it} policy gradient {e}: raw norm-wise error {r['raw']:.2e}, "
                  f"{r['ambiguous']} relu-ambiguous units ({r['flipped']} flipped), residual {r['residual']:.2e}")
            total_amb += r["ambiguous"]
            assert r["ambiguous"] < 0.35 * p0.size, "too many ambiguous directions for a meaningful projection"
    assert np.any(S["policy"].parameters() != p0)
    s = S["tr"].stats()
    assert s["env_steps"] == iters * n * T and s["reward_sum"] + s["episodes"] == s["env_steps"]
    for k in ("tr", "env", "value", "policy"):   # (a shared-trunk value model goes before its trunk's owner)
        S[k].close()


def test_env_sharding_is_rank_invariant_on_one_gpu(D, ctx):
    """'results independent of n'. Two shards of n/2 environments (env_offset 0
    and n/2, the two ranks of a 2-GPU run) against one shard of n on the same global env ids, all
    free-running on Philox streams keyed by the GLOBAL env id: rollout records bit-identical,
    per-shard flat gradients sum to the single-shard gradient within fp32 sum-order tolerance.
    (lr = 0 keeps the replicated parameters fixed, so the shards need no exchange here; the
    exchange itself is covered by tests/p2p_worker.py on 2+ GPUs.)"""
    n, T = 8192, 4

    def run(n_local, offset):
        policy = D.Model(ctx, D.fc_layers(PD, D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers(VD), 32)
        policy.init_parameters(11)
        value.init_parameters(12)
        env = D.Environment(ctx, n_local, seed=4321, env_offset=offset)
        tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=0.0, value_lr=0.0)
        outs = []
        for _ in range(2):
            tr.rollout()
            tr.learn()
            outs.append(dict(state=tr.read(D.F_REC_STATE), action=tr.read(D.F_REC_ACTION), done=tr.read(D.F_REC_DONE),
                             probs=tr.read(D.F_REC_PROBS), adv=tr.read(D.F_ADVANTAGE),
                             pg=tr.read(D.F_POLICY_GRAD_LOG).astype(np.float64), vg=tr.read(D.F_VALUE_GRAD).astype(np.float64),
                             live=env.state()))
        tr.close(); env.close(); policy.close(); value.close()
        return outs
    whole, a, b = run(n, 0), run(n // 2, 0), run(n // 2, n // 2)
    for it in range(2):
        w = whole[it]
        for key, axis in (("state", 2), ("action", 1), ("done", 1), ("probs", 1), ("adv", 1), ("live", 1)):
            both = np.concatenate([a[it][key], b[it][key]], axis=axis)
            assert np.array_equal(both, w[key]), f"iteration {it}: {key} depends on the sharding"
        close(a[it]["pg"] + b[it]["pg"], w["pg"], rtol=1e-5, what="policy gradient, 2 shards vs 1")
        close(a[it]["vg"] + b[it]["vg"], w["vg"], rtol=1e-5, what="value gradient, 2 shards vs 1")


def test_eval_after_graph_replay_uses_current_weights(D, ctx):
    """The learn phase replays as a CUDA graph from the third learn() on; Model.eval and
    eval_argmax (layered kernels with a transposed-weight cache) must see the replayed updates."""
    n, T = 2048, 4
    policy = D.Model(ctx, D.fc_layers(PD, D.SOFTMAX), 32)
    value = D.Model(ctx, D.fc_layers(VD), 32)
    policy.init_parameters(5)
    value.init_parameters(6)
    env = D.Environment(ctx, n, seed=3)
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-2 / (n * T), value_lr=1e-2 / (n * T))
    x = (np.random.default_rng(0).integers(0, 9, (64, 32)) / 8.0).astype(np.float32)
    for it in range(6):
        tr.iterate(1)
        fresh_p = D.Model(ctx, D.fc_layers(PD, D.SOFTMAX), 32)
        fresh_v = D.Model(ctx, D.fc_layers(VD), 32)
        fresh_p.set_parameters(policy.parameters())
        fresh_v.set_parameters(value.parameters())
        assert np.array_equal(policy.eval(x), fresh_p.eval(x)), f"policy eval is stale after iteration {it}"
        assert np.array_equal(value.eval(x), fresh_v.eval(x)), f"value eval is stale after iteration {it}"
        e1, e2 = D.Environment(ctx, 256, seed=9), D.Environment(ctx, 256, seed=9)
        assert D.eval_argmax(ctx, e1, policy, 1) == D.eval_argmax(ctx, e2, fresh_p, 1)
        for o in (fresh_p, fresh_v, e1, e2):
            o.close()
    tr.close(); env.close(); policy.close(); value.close()


@pytest.mark.parametrize("n,T,cap", [(65536, 4, 0), (3000, 4, 5), (2048, 8, 3), (777, 1, 2)])
def test_gae_local_end_rows_equal_the_prepass_launch(D, ctx, n, T, cap):
    """V(end) of the rows that end a trajectory: the GAE kernel's pipelines evaluate the end rows of their OWN
    tiles in compacted passes (list carried from tile to tile; default for large batches) -- same values, bit for
    bit, as the separate fused_vend_kernel launch (every row goes through the same row-wise GEMM chain), hence
    bit-identical advantages. cap: CTA cap = several tiles per pipeline at small sizes."""
    def run(mode):
        policy = D.Model(ctx, D.fc_layers(PD, D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers(VD), 32)
        policy.init_parameters(5)
        value.set_parameters((np.random.default_rng(6).standard_normal(value.n_params) * 0.1).astype(np.float32))
        env = D.Environment(ctx, n, seed=3)
        tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-3 / (n * T), value_lr=1e-3 / (n * T))
        if cap:
            D._lib.check(D._lib.lib.dfrl_debug_set_fused_ctas(tr.h, cap))
        D._lib.check(D._lib.lib.dfrl_debug_set_vend(tr.h, mode))
        out = []
        for _ in range(2):
            tr.rollout()
            tr.learn(D.PHASE_VALUE | D.PHASE_ADVANTAGE)
            out.append((tr.read(D.F_ADVANTAGE).copy(), tr.read(D.F_REC_DONE).copy()))
            tr.learn(D.PHASE_POLICY)
        tr.close(); env.close(); policy.close(); value.close()
        return out
    local, launch = run(1), run(2)
    for it in range(2):
        assert np.array_equal(local[it][1], launch[it][1])
        assert local[it][1].any() or T == 1
        assert np.array_equal(local[it][0], launch[it][0]), f"iteration {it}: advantages differ"
        assert np.all(np.isfinite(local[it][0])) and np.abs(local[it][0]).max() > 0


def test_kl_ppo_learn_phase_replays_as_a_graph(D, ctx):
    """KL-PPO (kl_ppo_learner, policy_gradient.h:310-335): critic step / GAE on the fused kernels, the k policy
    steps on the layered kernels, beta adapted on the device between the steps -- no host round trip, so the
    learn phase is captured as a CUDA graph from the third learn() on. Same parameters and beta, bit for bit,
    as the launch-by-launch path (learn_phases never replays a graph)."""
    n, T = 1024, 4
    def run(graphed):
        policy = D.Model(ctx, D.fc_layers(PD, D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers(VD), 32)
        policy.init_parameters(5)
        value.init_parameters(6)
        env = D.Environment(ctx, n, seed=3)
        lr = 1e-2 / (n * T)
        tr = D.Trainer(ctx, env, policy, value, algo=D.KL_PPO, work=T, policy_lr=lr, value_lr=lr, policy_wd=1e-5,
                       kl_target=1e-3, kl_beta0=0.05)
        out = []
        launches = []
        for it in range(7):
            tr.rollout()
            l0 = ctx.launches()
            if graphed:
                tr.learn()
            else:
                tr.learn(D.PHASE_VALUE | D.PHASE_ADVANTAGE)
                tr.learn(D.PHASE_POLICY)
            launches.append(ctx.launches() - l0)
            out.append((policy.parameters().copy(), value.parameters().copy(), tr.stats()["kl_beta"]))
        cov = tr.fused_coverage()
        tr.close(); env.close(); policy.close(); value.close()
        return out, launches, cov
    (g, gl, cov), (l, ll, _) = run(True), run(False)
    assert cov & D.FUSED_CRITIC
    betas = [x[2] for x in g]
    assert len(set(betas)) > 1, betas          # beta really adapts at this target
    for it in range(7):
        assert np.array_equal(g[it][0], l[it][0]) and np.array_equal(g[it][1], l[it][1]), f"iteration {it}"
        assert g[it][2] == l[it][2], (it, g[it][2], l[it][2])
    assert gl[-1] == ll[-1]                    # the replayed graph stands for the same kernels


def test_conv_table_path_is_deterministic_and_agrees_with_the_tensor_core_path(D, ctx, monkeypatch):
    """conv1d policy, 16 384 envs x 4 (the table path's default range): two runs bit-identical (the histogram of dY
    accumulates in fixed point: integer atomics commute), free-running rollouts identical to the tensor-core
    kernels' until probabilities differ in the last bits, gradients of the first policy step within 1e-4 of theirs,
    and the learn phase replays as a CUDA graph."""
    n, T = 16384, 4
    def run(table, iters):
        monkeypatch.setenv("DFRL_CONV_TABLE", str(table))
        policy = D.Model(ctx, D.conv_layers([4, 128, 64, 1], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 32, 1]), 32)
        policy.init_parameters(21)
        value.init_parameters(22)
        env = D.Environment(ctx, n, seed=5)
        tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-4 * 32 / (n * T), value_lr=1e-5 * 32 / (n * T))
        assert tr.fused_covers_iteration()
        out = []
        for it in range(iters):
            tr.rollout()
            acts, probs = tr.read(D.F_REC_ACTION).copy(), tr.read(D.F_REC_PROBS).copy()
            tr.learn()
            out.append(dict(acts=acts, probs=probs, g=tr.read(D.F_POLICY_GRAD_LOG).copy(), p=policy.parameters().copy(),
                            adv=tr.read(D.F_ADVANTAGE).copy()))
        tr.close(); env.close(); policy.close(); value.close()
        return out
    a, b, c = run(1, 5), run(1, 5), run(0, 1)
    for it in range(5):   # iterations 3.. replay the captured graph
        for key in ("acts", "probs", "g", "p", "adv"):
            assert np.array_equal(a[it][key], b[it][key]), (it, key)
    assert np.all(np.isfinite(a[-1]["p"])) and np.any(a[-1]["p"] != a[0]["p"])
    # against the tensor-core kernels: same first rollout up to rounding of the probabilities, same gradients to 1e-4
    assert np.abs(a[0]["probs"] - c[0]["probs"]).max() < 1e-5
    same = np.mean(a[0]["acts"] == c[0]["acts"])
    assert same > 0.9999, same
    if same == 1.0:
        for e in range(4):
            close(a[0]["g"][e].astype(np.float64), c[0]["g"][e].astype(np.float64), rtol=1e-4, what=f"policy gradient {e}: table vs tensor cores")


def test_conv_table_reuse_follows_the_parameters(D, ctx, monkeypatch):
    """PPO's first policy step sees the parameters of the rollout, so it takes the rollout's logit table instead of
    computing its own (fused.cu: tbl_version == the policy's parameter version). The reuse must stop the moment the
    parameters are not the rollout's: parameters replaced between rollout and learn, a second learn() without a
    rollout -- also across the CUDA-graph replay of the learn phase, which is captured under the key
    dfrl_fused_learn_key() and captured again when the key changes. Bit-identical to a trainer whose every step
    computes its table (DFRL_TABLE_REUSE=0) through the same schedule."""
    n, T = 16384, 4
    def run(reuse):
        monkeypatch.setenv("DFRL_CONV_TABLE", "1")
        monkeypatch.setenv("DFRL_TABLE_REUSE", str(reuse))
        policy = D.Model(ctx, D.conv_layers([4, 128, 64, 1], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 32, 1]), 32)
        policy.init_parameters(31)
        value.init_parameters(32)
        env = D.Environment(ctx, n, seed=7)
        tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-4 * 32 / (n * T), value_lr=1e-5 * 32 / (n * T))
        out = []
        for it in range(9):
            if it != 6:                       # iteration 6: learn() again on the same records, new parameters
                tr.rollout()
            if it == 4:                       # parameters replaced between rollout and learn
                policy.set_parameters((policy.parameters() * np.float32(1.02)).astype(np.float32))
            tr.learn()
            out.append(dict(g=tr.read(D.F_POLICY_GRAD_LOG).copy(), p=policy.parameters().copy()))
        tr.close(); env.close(); policy.close(); value.close()
        return out
    a, b = run(1), run(0)
    for it in range(9):
        for key in ("g", "p"):
            assert np.array_equal(a[it][key], b[it][key]), (it, key)
    assert np.any(a[4]["g"][0] != a[3]["g"][0])


def test_set_rates_on_a_graph_replaying_learner(D, ctx):
    """optimizer::set_rate (nn.h:592) after the learn phase has been captured as a CUDA graph: the new
    rates take effect at the next learn() (rates are kernel arguments: the graph is re-captured), the
    momentum state survives. Checked against a second trainer that is driven launch by launch
    (learn_phases never replays a graph) through the same schedule."""
    n, T = 1024, 4
    def run(graphed):
        policy = D.Model(ctx, D.fc_layers(PD, D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers(VD), 32)
        policy.init_parameters(5)
        value.init_parameters(6)
        env = D.Environment(ctx, n, seed=3)
        lr = 1e-2 / (n * T)
        tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=lr, value_lr=lr,
                       policy_opt=D.MOMENTUM, value_opt=D.MOMENTUM)
        out = []
        for it in range(8):
            if it == 5:
                tr.set_rates(3 * lr, 0.5 * lr)
            tr.rollout()
            if graphed:
                tr.learn()
            else:
                tr.learn(D.PHASE_VALUE | D.PHASE_ADVANTAGE)
                tr.learn(D.PHASE_POLICY)
            out.append((policy.parameters().copy(), value.parameters().copy()))
        tr.close(); env.close(); policy.close(); value.close()
        return out
    g, l = run(True), run(False)
    assert not np.array_equal(g[4][0], g[5][0])
    for it in range(8):
        assert np.array_equal(g[it][0], l[it][0]) and np.array_equal(g[it][1], l[it][1]), f"iteration {it}"


@pytest.mark.parametrize("nets,iters", [("reference", 1000), ("c2_fused_he_init", 3000)])
def test_ppo_learns_bin_packing(D, ctx, nets, iters):
    """The reference's only own test is the reward of the argmax policy climbing towards 26.55
    (ppo_training.cc:67-81, deep.log; random placement ~ 11.6, the untrained argmax policy ~ 3). PPO
    at 4096 envs with the reference's nets / rates (layered path) and with the fused C2 nets
    (He-scaled dense init: see tools/convergence.py) must clear 20 -- the unmodified reference
    reaches 21.7 after 2500 rounds of 8 envs (profiles/r02_convergence_ref_cpu.csv)."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import convergence
    c = convergence.curve(ctx, nets, iters=iters, every=iters // 4)
    print(nets, c)
    assert c[0][2] < 12 and c[-1][2] > 20, c
