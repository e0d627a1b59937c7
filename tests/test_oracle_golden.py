"""CPU suite: oracle/dfrl_oracle.c (our restatement) against the golden fixtures that
tests/golden/make_golden.py generated from the UNMODIFIED reference (oracle/_ref).

Bit-exact for integer/index work (RNG, sampling, env transitions); 1e-4 relative for fp32."""
import os

import numpy as np
import pytest

import refcases
from refcases import close
from oracle import orc

U = dict(np.load(os.path.join(refcases.GOLDEN_DIR, "units.npz")))


def test_minstd_engine_bit_exact():
    for seed, key in [(1, "engine_seed1"), (1234, "engine_seed1234")]:
        g = orc.Minstd(seed)
        got = np.array([g.next() for _ in range(16)], dtype=np.uint32)
        assert np.array_equal(got, U[key])


@pytest.mark.parametrize("wkey,skey,seed", [("disc_w", "disc_samples_seed9", 9),
                                            ("disc_w2", "disc_samples2_seed77", 77)])
def test_discrete_distribution_bit_exact(wkey, skey, seed):
    # std::discrete_distribution on the seeded engine == orc_discrete(w, generate_canonical)
    g = orc.Minstd(seed)
    got = np.array([orc.discrete(U[wkey], g.canonical()) for _ in range(256)], dtype=np.int32)
    assert np.array_equal(got, U[skey])


def test_argmax_first_max():
    assert orc.argmax(U["argmax_ties"]) == int(U["argmax_ties_idx"]) == 1


@pytest.mark.parametrize("name,n_in,n_out", [("dense_32_64", 32, 64), ("dense_64_1", 64, 1),
                                             ("dense_5_3", 5, 3), ("conv_4_16", 4, 16),
                                             ("conv_16_1", 16, 1)])
def test_dense_and_conv_layers(name, n_in, n_out):
    p, x, dy = U[f"{name}_p"], U[f"{name}_x"], U[f"{name}_dy"]
    rows = x.shape[0]
    # conv1d_1 == dense over (rows * points, channels) (nn.h:127-147)
    x2, dy2 = x.reshape(-1, n_in), dy.reshape(-1, n_out)
    close(orc.dense_forward(p, n_in, n_out, x2).reshape(rows, -1), U[f"{name}_y"], what="fwd")
    close(orc.dense_backward(p, n_in, n_out, dy2).reshape(rows, -1), U[f"{name}_dx"], what="bwd")
    close(orc.dense_gradient(n_in, n_out, x2, dy2), U[f"{name}_g"], what="grad")


def test_activations():
    close(orc.relu_forward(U["relu_x"]), U["relu_y"], what="relu fwd")
    close(orc.relu_backward(U["relu_x"], U["relu_dy"]), U["relu_dx"], what="relu bwd")
    close(orc.softmax_forward(U["softmax_x"]), U["softmax_y"], what="softmax fwd")
    close(orc.softmax_backward(U["softmax_x"], U["softmax_dy"]), U["softmax_dx"], what="softmax bwd")
    # softmax_cross_entropy_layer: same forward, identity backward (nn.h:424-431)
    close(orc.softmax_forward(U["softmax_ce_x"]), U["softmax_ce_y"], what="softmax_ce fwd")
    assert np.array_equal(U["softmax_ce_dx"], U["softmax_ce_dy"])


@pytest.mark.parametrize("name", ["mlp_c2_policy", "mlp_value", "mlp_conv_policy", "mlp_pg_policy"])
def test_models(name):
    net = refcases.orc_net_from_layers(U[f"{name}_layers"])
    g, out = orc.net_forward_gradient(net, U[f"{name}_p"], U[f"{name}_x"], U[f"{name}_dy"])
    close(out, U[f"{name}_out"], what="model out")
    close(g, U[f"{name}_g"], what="model grad")
    close(orc.net_eval(net, U[f"{name}_p"], U[f"{name}_x"]), U[f"{name}_out"], what="eval")
    # the double-accumulating variant agrees too
    g64, _ = orc.net_forward_gradient(net, U[f"{name}_p"], U[f"{name}_x"], U[f"{name}_dy"], f64=True)
    close(g64, U[f"{name}_g"], what="model grad f64")


def test_loss_gradients():
    P, PO, ch, adv = U["loss_p"], U["loss_pold"], U["loss_choice"], U["loss_adv"]
    close(orc.loss_grad(orc.LOSS_SOFTMAX_LOG, P, ch, adv), U["loss_softmax_log"], what="log")
    po_sel = PO[np.arange(40), ch]
    close(orc.loss_grad(orc.LOSS_CLIPPED, P, ch, adv, po_sel), U["loss_clipped"], what="clip")
    close(orc.loss_grad(orc.LOSS_KL, P, ch, adv, PO, beta=1.0), U["loss_kl_beta1"], what="kl")
    assert orc.kl_next_beta(P, PO, 1e-9, 1.0) == pytest.approx(float(U["loss_kl_beta_next"]))
    close(orc.loss_grad(orc.LOSS_KL, P, ch, adv, P, beta=0.05), U["loss_kl_same"], what="kl same")
    assert orc.kl_next_beta(P, P, 1e-9, 0.05) == pytest.approx(float(U["loss_kl_same_beta_next"]))


def test_ppo_clip_known_answers():
    # SURVEY.md appendix A (hand-checked against rl.h:54-74): gradient non-zero when clipped
    po = np.float32(0.125)
    for pa, A, want in [(0.125, 1.5, -12), (0.125, -1.5, 12), (0.2, 1.5, -9), (0.2, -1.5, 12),
                        (0.05, 1.5, -12), (0.05, -1.5, 24)]:
        p = np.full((1, 8), (1 - pa) / 7, np.float32)
        p[0, 3] = pa
        g = orc.loss_grad(orc.LOSS_CLIPPED, p, [3], [A], [po])
        assert g[0, 3] == pytest.approx(want, rel=1e-6)
        assert np.count_nonzero(g) == 1


@pytest.mark.parametrize("key,kind,wd", [("opt_sgd", orc.SGD, 0.0), ("opt_sgd_wd", orc.SGD, 1e-3),
                                         ("opt_momentum", orc.MOMENTUM, 0.0), ("opt_adam", orc.ADAM, 0.0)])
def test_optimizers(key, kind, wd):
    p = U["opt_p0"].copy()
    state = np.zeros(max(1, orc.opt_state_size(kind, p.size)), np.float32)
    t = 1.0
    for k, g in enumerate(U["opt_grads"]):
        orc.opt_step(kind, p, g, state, 1e-2, wd, 0.9, 0.999, t)
        t += 1.0
        close(p, U[key][k], what=f"{key} step {k}")


@pytest.mark.parametrize("akey,skey", [("env_forced_actions", "env_forced_steps"),
                                       ("env_rr_actions", "env_rr_steps")])
def test_env_dynamics_bit_exact(akey, skey):
    steps = U[skey]
    cfg = orc.env_cfg()
    first = refcases.item_code(steps["sitem"][0])
    st = orc.env_reset_all(cfg, 1, [first])
    n_done = 0
    for r in steps:
        assert np.array_equal(st[:, 0], refcases.planes(r["sbins"], r["sitem"])), "start state"
        done, term = orc.env_step(cfg, st, [r["action"]], [refcases.item_code(r["item_after"])], True)
        assert done[0] == r["done"]
        assert (0.0 if done[0] else 1.0) == r["reward"]
        assert np.array_equal(term[:2 * 8, 0], refcases.planes(r["ebins"], r["eitem"])[:16]), "end bins"
        if r["done"]:
            n_done += 1
            # terminal state keeps the item that overflowed (bin_packing.h:59-61)
            assert np.array_equal(term[16:, 0], np.array(r["sitem"], np.int8))
            assert np.all(st[:16, 0] == 8)
        else:
            # non-terminal end_state = env.view() after apply, new item included (rl.h:336)
            assert np.array_equal(st[:, 0], refcases.planes(r["ebins"], r["eitem"]))
    assert n_done >= 5


def test_obs_encode_exact():
    steps = U["env_forced_steps"][:50]
    st = refcases.planes(steps["sbins"], steps["sitem"])
    obs = orc.obs_encode(st, 8)
    want = np.zeros((50, 8, 4), np.float32)
    want[:, :, 0] = steps["sbins"][:, :, 0] / 8.0
    want[:, :, 1] = steps["sbins"][:, :, 1] / 8.0
    want[:, :, 2] = steps["sitem"][:, None, 0] / 8.0
    want[:, :, 3] = steps["sitem"][:, None, 1] / 8.0
    assert np.array_equal(obs, want.reshape(50, 32))


def test_item_stream_is_seeded_bernoulli():
    # the env's item draws are bernoulli(0.4) on the global engine (bin_packing.h:50, 76-81):
    # with a forced policy no other draws happen, so the tape must equal the restated stream.
    steps = U["env_rr_steps"]
    g = orc.Minstd(3)
    want = [g.bernoulli(0.4)]  # constructor draw
    for _ in steps:
        want.append(g.bernoulli(0.4))
    got = [refcases.item_code(steps["sitem"][0])] + [refcases.item_code(r["item_after"]) for r in steps]
    assert list(map(int, got)) == want


# ---------------------------------------------------------------- whole learner iterations ---

@pytest.mark.parametrize("name", refcases.case_names())
def test_learner_against_reference_trace(name):
    c = refcases.load_case(name)
    algo, n, work, iters = int(c["algo"]), int(c["n_envs"]), int(c["work"]), int(c["iters"])
    pnet = refcases.orc_net_from_layers(c["policy_layers"])
    vnet = refcases.orc_net_from_layers(c["value_layers"]) if len(c["value_layers"]) else None
    tcfg = orc.train_cfg(algo, work, policy_opt=int(c["popt"]), value_opt=int(c["vopt"]),
                         policy_lr=float(c["plr"]), value_lr=float(c["vlr"]), policy_wd=float(c["pwd"]))
    ecfg = orc.env_cfg()
    lr = orc.Learner(tcfg, ecfg, pnet, c["pparams0"], vnet, c["vparams0"] if vnet else None)
    state = refcases.initial_state(c["steps"], n)
    pg_i = vg_i = 0
    epochs = 1 if algo in (orc.REINFORCE, orc.ACTOR_CRITIC) else 4
    for it in range(iters):
        rec = refcases.records_from_steps(c["steps"], it, n)
        L = rec["L"]
        # rollout: teacher-forced actions + the reference's item tape; transitions bit-exact,
        # policy outputs (p_old) within tolerance
        if algo != orc.REINFORCE:
            ro = orc.rollout(ecfg, state, pnet, lr.pparams, L, 2, rec["items"], forced=rec["action"])
            assert np.array_equal(ro["state"], rec["rec_state"])
            assert np.array_equal(ro["done"], rec["done"])
            assert np.array_equal(state, rec["final_state"])
            close(ro["probs"], rec["p_old"], what=f"{name} iter {it} p_old")
        out = lr.learn(rec["rec_state"], rec["final_state"], rec["action"], rec["done"], rec["p_old"],
                       rec["len"] if algo == orc.REINFORCE else None)
        adv_ref = refcases.adv_from_rows(c["rows"], it, n, L)
        mask = np.arange(L)[:, None] < rec["len"][None, :]
        close(out["adv"][mask], adv_ref[mask], what=f"{name} iter {it} advantages")
        if vnet is not None:
            close(out["value_grad"], c["value_grads"][vg_i], what=f"{name} iter {it} value grad")
            vg_i += 1
        for e in range(epochs):
            close(out["policy_grads"][e], c["policy_grads"][pg_i], what=f"{name} iter {it} policy grad {e}")
            close(lr.pparams if e == epochs - 1 else c["policy_params_log"][pg_i], c["policy_params_log"][pg_i],
                  what="params")
            pg_i += 1
    close(lr.pparams, c["pparams_final"], what=f"{name} final policy params")
    if vnet is not None:
        close(lr.vparams, c["vparams_final"], what=f"{name} final value params")
    assert pg_i == len(c["policy_grads"])


def test_reinforce_returns_known_answer():
    # rewards (1,2,3,0) cannot occur in bin packing (rewards are 1..1,0), so
    # the KAT is restated for the env's reward pattern: one frozen trajectory of 4 steps,
    # gamma = 0.5: forward pass writes backward -> G = [0.875, 1.75, 1.5, 1.0]
    done = np.array([[0], [0], [0], [1]], np.uint8)
    g, acc = orc.returns(done, [4], 0.5)
    assert np.allclose(g[:, 0], [0.875, 1.75, 1.5, 1.0])
    assert acc[0] == pytest.approx(0.875) and acc[1] == 1


def test_gae_known_answer():
    # V == 2 everywhere, gamma 0.5, lambda 0.95; frozen trajectory of 4 then an open one of 4
    T = 8
    done = np.zeros((T, 1), np.uint8)
    done[3, 0] = 1
    v = np.full((T, 1), 2.0, np.float32)
    tg, adv = orc.gae(done, v, v, 0.5, 0.95)
    # targets: r + 0.5 * 2 (unmasked even at the terminal, quirk 6)
    assert np.allclose(tg[:, 0], [2, 2, 2, 1, 2, 2, 2, 2])
    c = 0.5 * 0.95
    d = np.array([0, 0, 0, -2.0, 0, 0, 0, 0])  # delta: r + g*Vn - V ; terminal: 0 + 0 - 2
    want = np.zeros(T)
    want[3] = d[3]
    for t in (2, 1, 0):
        want[t] = d[t] + c * want[t + 1]
    assert np.allclose(adv[:4, 0], want[:4], atol=1e-6)
    assert np.allclose(adv[4:, 0], 0, atol=1e-6)
