"""GPU parity suite (-m gpu): libdfrl_b200.so through its C ABI against the CPU oracle
(oracle/dfrl_oracle.c) on identical seeded inputs and against the golden fixtures generated
from the unmodified reference.  Bit-exact for integer / byte / index work; 1e-4 relative for
fp32 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import refcases
from refcases import close

pytestmark = pytest.mark.gpu

U = dict(np.load(os.path.join(refcases.GOLDEN_DIR, "units.npz")))




def test_device_is_blackwell(ctx):
    info = ctx.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] >= 100


# ------------------------------------------------------------------------- K1 environment ---

@pytest.mark.parametrize("n,B", [(1, 8), (7, 8), (64, 8), (1000, 8), (1003, 8), (256, 16), (33, 5),
                                 (128, 32)])
def test_env_step_bit_exact(D, ctx, orc, n, B):
    rng = np.random.default_rng(n * 131 + B)
    steps = 60
    tape = rng.integers(0, 2, (n, steps + 1)).astype(np.uint8)
    env = D.Environment(ctx, n, n_bins=B)
    env.load_item_tape(tape)
    env.reset()
    cfg = orc.env_cfg(B)
    st = orc.env_reset_all(cfg, n, tape[:, 0])
    assert np.array_equal(env.state(), st)
    dones = 0
    for t in range(steps):
        # bias towards a few bins so that overflows (episode ends) happen often
        a = rng.integers(0, min(B, 3), n).astype(np.uint8) if t % 2 else rng.integers(0, B, n).astype(np.uint8)
        d_gpu, term_gpu = env.step(a, want_terminal=True)
        d_cpu, term_cpu = orc.env_step(cfg, st, a, tape[:, t + 1], want_terminal=True)
        assert np.array_equal(d_gpu, d_cpu), f"done mismatch at step {t}"
        assert np.array_equal(term_gpu, term_cpu), f"terminal state mismatch at step {t}"
        assert np.array_equal(env.state(), st), f"state mismatch at step {t}"
        dones += int(d_cpu.sum())
    assert dones > 0
    assert np.array_equal(env.obs(), orc.obs_encode(st, B))
    env.close()


def test_env_matches_reference_trace(D, ctx):
    # golden: bp::environment + bp::agent stepped by the reference with forced actions
    for akey, skey in [("env_forced_actions", "env_forced_steps"), ("env_rr_actions", "env_rr_steps")]:
        steps = U[skey]
        tape = np.concatenate([[refcases.item_code(steps["sitem"][0])],
                               refcases.item_code(steps["item_after"])]).astype(np.uint8)[None, :]
        env = D.Environment(ctx, 1)
        env.load_item_tape(tape)
        env.reset()
        for r in steps:
            assert np.array_equal(env.state()[:, 0], refcases.planes(r["sbins"], r["sitem"]))
            d, term = env.step(np.array([r["action"]], np.uint8), want_terminal=True)
            assert d[0] == r["done"]
            assert np.array_equal(term[:16, 0], refcases.planes(r["ebins"], r["eitem"])[:16])
        env.close()


def test_env_philox_items_are_bernoulli_04_and_rank_invariant(D, ctx):
    n = 1 << 16
    env = D.Environment(ctx, n, seed=99)
    st = env.state()
    frac = np.mean(st[16] == 4)
    assert abs(frac - 0.4) < 0.01
    # sharding by env offset reproduces the same global stream (multi-GPU invariance)
    half = D.Environment(ctx, n // 2, seed=99, env_offset=n // 2)
    assert np.array_equal(half.state(), st[:, n // 2:])
    env.close()
    half.close()


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_heuristic_react_matches_oracle(D, ctx, orc, kind):
    n, B = 512, 8
    rng = np.random.default_rng(kind)
    env = D.Environment(ctx, n, seed=5)
    # random reachable-looking states
    st = env.state()
    st[:16] = rng.integers(0, 9, (16, n))
    st[17] = 2
    st[16] = np.where(rng.random(n) < 0.4, 4, 1)
    env.set_state(st)
    got = env.heuristic_react(kind)
    cfg = orc.env_cfg(B)
    if kind == 0:
        assert got.min() >= 0 and got.max() < B and len(np.unique(got)) == B
    else:
        want = np.array([orc.heuristic_react(cfg, st, i, kind) for i in range(n)], np.uint8)
        assert np.array_equal(got, want)
    env.close()


def test_heuristic_levels_match_reference_logs(D, ctx):
    # reference logs: min-waste 26.553 +- 0.009 (minwaste.log), first-fit ~25.83, best-fit ~25.81,
    # random ~11.6
    for kind, lo, hi in [(3, 26.45, 26.65), (1, 25.7, 25.95), (2, 25.7, 25.95), (0, 11.2, 12.0)]:
        env = D.Environment(ctx, 4096, seed=7 + kind)
        total, steps = env.heuristic_play(kind, 8)
        mean = total / (4096 * 8)
        assert lo < mean < hi, (kind, mean)
        assert steps == total + 4096 * 8  # every episode ends with exactly one zero-reward step
        env.close()


# ------------------------------------------------------------------------- K2/K6 layers -----

def _dense_case(D, ctx, orc, rows, n_in, n_out, seed, relu=False):
    rng = np.random.default_rng(seed)
    p = (rng.standard_normal((n_in + 1) * n_out) * 0.3).astype(np.float32)
    x = rng.standard_normal((rows, n_in)).astype(np.float32)
    dy = rng.standard_normal((rows, n_out)).astype(np.float32)
    lib, chk = D._lib.lib, D._lib.check
    dp, dx, ddy = ctx.to_device(p), ctx.to_device(x), ctx.to_device(dy)
    y = ctx.empty((rows, n_out), np.float32)
    chk(lib.dfrl_dense_forward(ctx.h, dp.p, n_in, n_out, dx.p, rows, y.p, 1 if relu else 0))
    want = orc.dense_forward(p, n_in, n_out, x)
    if relu:
        want = np.maximum(want, 0)
    close(y.get(), want, what=f"fwd {rows}x{n_in}->{n_out}")
    gx = ctx.empty((rows, n_in), np.float32)
    chk(lib.dfrl_dense_backward(ctx.h, dp.p, n_in, n_out, ddy.p, rows, None, gx.p))
    close(gx.get(), orc.dense_backward(p, n_in, n_out, dy), what="bwd")
    # fused relu mask of the preceding activation
    chk(lib.dfrl_dense_backward(ctx.h, dp.p, n_in, n_out, ddy.p, rows, dx.p, gx.p))
    close(gx.get(), orc.dense_backward(p, n_in, n_out, dy) * (x > 0), what="bwd masked")
    g = ctx.empty(((n_in + 1) * n_out,), np.float32)
    chk(lib.dfrl_dense_gradient(ctx.h, n_in, n_out, dx.p, ddy.p, rows, g.p, 0))
    close(g.get(), orc.dense_gradient(n_in, n_out, x, dy, f64=True), what="grad")
    for a in (dp, dx, ddy, y, gx, g):
        a.free()


@pytest.mark.parametrize("rows,n_in,n_out", [(1, 32, 64), (37, 32, 64), (1000, 64, 64), (333, 64, 8),
                                             (129, 64, 1), (4, 5, 3), (777, 4, 128), (513, 128, 64),
                                             (260, 256, 256), (100, 128, 32), (64, 64, 32), (50, 32, 16),
                                             # >= 4096 rows: the tcgen05 GEMMs of gemm_umma.cu (C5 shapes,
                                             # ragged last tile, M = 64 / 128 / 2 x 128 output blocks)
                                             (5000, 128, 256), (4100, 256, 256), (6001, 256, 32),
                                             (4096, 32, 64), (4500, 64, 16), (8192, 256, 128),
                                             (4097, 48, 64), (4200, 64, 8)])
def test_dense_layer_vs_oracle(D, ctx, orc, rows, n_in, n_out):
    _dense_case(D, ctx, orc, rows, n_in, n_out, rows + n_in + n_out)
    _dense_case(D, ctx, orc, rows, n_in, n_out, 3, relu=True)


def test_dense_layers_vs_reference_golden(D, ctx):
    lib, chk = D._lib.lib, D._lib.check
    for name, n_in, n_out in [("dense_32_64", 32, 64), ("dense_64_1", 64, 1), ("dense_5_3", 5, 3),
                              ("conv_4_16", 4, 16), ("conv_16_1", 16, 1)]:
        p, x, dy = U[f"{name}_p"], U[f"{name}_x"], U[f"{name}_dy"]
        rows0 = x.shape[0]
        x2, dy2 = x.reshape(-1, n_in), dy.reshape(-1, n_out)
        rows = x2.shape[0]
        dp, dx, ddy = ctx.to_device(p), ctx.to_device(x2), ctx.to_device(dy2)
        y = ctx.empty((rows, n_out), np.float32)
        gx = ctx.empty((rows, n_in), np.float32)
        g = ctx.empty((p.size,), np.float32)
        chk(lib.dfrl_dense_forward(ctx.h, dp.p, n_in, n_out, dx.p, rows, y.p, 0))
        chk(lib.dfrl_dense_backward(ctx.h, dp.p, n_in, n_out, ddy.p, rows, None, gx.p))
        chk(lib.dfrl_dense_gradient(ctx.h, n_in, n_out, dx.p, ddy.p, rows, g.p, 0))
        close(y.get().reshape(rows0, -1), U[f"{name}_y"], what=name + " y")
        close(gx.get().reshape(rows0, -1), U[f"{name}_dx"], what=name + " dx")
        close(g.get(), U[f"{name}_g"], what=name + " g")


def test_activations_vs_reference_golden(D, ctx):
    lib, chk = D._lib.lib, D._lib.check
    x, dy = ctx.to_device(U["relu_x"]), ctx.to_device(U["relu_dy"])
    y, dx = ctx.empty(U["relu_x"].shape, np.float32), ctx.empty(U["relu_x"].shape, np.float32)
    chk(lib.dfrl_relu_forward(ctx.h, x.p, x.size, y.p))
    chk(lib.dfrl_relu_backward(ctx.h, x.p, dy.p, x.size, dx.p))
    close(y.get(), U["relu_y"], what="relu y")
    close(dx.get(), U["relu_dx"], what="relu dx")
    x, dy = ctx.to_device(U["softmax_x"]), ctx.to_device(U["softmax_dy"])
    chk(lib.dfrl_softmax_forward(ctx.h, x.p, 23, 8, y.p))
    chk(lib.dfrl_softmax_backward(ctx.h, x.p, dy.p, 23, 8, dx.p))
    close(y.get(), U["softmax_y"], what="softmax y")
    close(dx.get(), U["softmax_dx"], what="softmax dx")


@pytest.mark.parametrize("name", ["mlp_c2_policy", "mlp_value", "mlp_conv_policy", "mlp_pg_policy"])
def test_models_vs_reference_golden(D, ctx, name):
    m = D.Model(ctx, U[f"{name}_layers"], U[f"{name}_x"].shape[1])
    assert m.n_params == U[f"{name}_p"].size
    m.set_parameters(U[f"{name}_p"])
    assert np.array_equal(m.parameters(), U[f"{name}_p"])
    close(m.eval(U[f"{name}_x"]), U[f"{name}_out"], what="eval")
    g, out = m.forward_gradient(U[f"{name}_x"], U[f"{name}_dy"])
    close(out, U[f"{name}_out"], what="out")
    close(g, U[f"{name}_g"], what="grad")
    m.close()


@pytest.mark.parametrize("rows", [1, 5000])
def test_c5_wide_model_vs_oracle(D, ctx, orc, rows):
    # C5: B = 32 bins, 128 -> 256 -> 256 -> 256 -> 32 + softmax
    layers = D.fc_layers([128, 256, 256, 256, 32], D.SOFTMAX)
    m = D.Model(ctx, layers, 128)
    rng = np.random.default_rng(rows)
    p = (rng.standard_normal(m.n_params) * 0.05).astype(np.float32)
    x = (rng.integers(0, 9, (rows, 128)) / 8.0).astype(np.float32)
    dy = rng.standard_normal((rows, 32)).astype(np.float32)
    # relu'(x) is discontinuous at 0: a pre-activation within rounding distance of zero takes
    # either sign under equally valid summation orders and moves a whole gradient row. Rows with
    # such a pre-activation (fp64 forward, margin 1e-5) get a zero upstream gradient so that the
    # comparison below is a strict 1e-4 one.
    h, off, dims = x.astype(np.float64), 0, [128, 256, 256, 256, 32]
    risky = np.zeros(rows, dtype=bool)
    for a, b in zip(dims[:-2], dims[1:-1]):
        W = p[off:off + a * b].reshape(b, a).astype(np.float64)
        bias = p[off + a * b:off + a * b + b].astype(np.float64)
        off += a * b + b
        pre = h @ W.T + bias
        risky |= (np.abs(pre) < 1e-5).any(axis=1)
        h = np.maximum(pre, 0.0)
    dy[risky] = 0.0
    m.set_parameters(p)
    net = orc.Net(layers, 128)
    assert net.param_count() == m.n_params == 172832
    g, out = m.forward_gradient(x, dy)
    g_want, out_want = orc.net_forward_gradient(net, p, x, dy, f64=True)
    close(out, out_want, what="c5 out")
    close(g, g_want, what="c5 grad")
    m.close()


def test_model_init_statistics(D, ctx):
    # nn.h:12-18: dense N(0, 0.01), conv1d He N(0, sqrt(2/in)), biases zero
    m = D.Model(ctx, D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
    m.init_parameters(5)
    p = m.parameters()
    w0 = p[:32 * 64]
    assert abs(w0.std() - 0.01) < 0.001 and abs(w0.mean()) < 0.001
    assert abs(float(U["init_dense_std"]) - 0.01) < 0.001
    assert np.all(p[32 * 64:32 * 64 + 64] == 0)
    c = D.Model(ctx, D.conv_layers([4, 128, 64, 1]), 32)
    c.init_parameters(5)
    pc = c.parameters()
    assert abs(pc[:512].std() - np.sqrt(2 / 4)) < 0.06
    assert abs(float(U["init_conv_std_l1"]) - np.sqrt(2 / 4)) < 0.06
    m.close()
    c.close()


# ------------------------------------------------------------------------- K3 sampling ------

def test_sampling_bit_exact_vs_oracle_and_reference(D, ctx, orc):
    lib, chk = D._lib.lib, D._lib.check
    rng = np.random.default_rng(3)
    rows = 4096
    probs = rng.random((rows, 8)).astype(np.float32) ** 3
    probs /= probs.sum(1, keepdims=True)
    u = rng.random(rows)
    u[:4] = [0.0, 1.0 - 2 ** -53, 0.5, 0.125]
    dp, du = ctx.to_device(probs), ctx.to_device(u)
    a = ctx.empty((rows,), np.uint8)
    ps = ctx.empty((rows,), np.float32)
    chk(lib.dfrl_sample(ctx.h, dp.p, rows, 8, du.p, a.p, ps.p))
    got = a.get()
    want = np.array([orc.discrete(probs[i], u[i]) for i in range(rows)], np.uint8)
    assert np.array_equal(got, want)
    assert np.array_equal(ps.get(), probs[np.arange(rows), got])
    # the reference's own draws: std::discrete_distribution on the seeded engine
    for wkey, skey, seed in [("disc_w", "disc_samples_seed9", 9), ("disc_w2", "disc_samples2_seed77", 77)]:
        g = orc.Minstd(seed)
        uu = np.array([g.canonical() for _ in range(256)])
        pw = np.tile(U[wkey], (256, 1)).astype(np.float32)
        dp2, du2 = ctx.to_device(pw), ctx.to_device(uu)
        a2 = ctx.empty((256,), np.uint8)
        chk(lib.dfrl_sample(ctx.h, dp2.p, 256, 8, du2.p, a2.p, None))
        assert np.array_equal(a2.get().astype(np.int32), U[skey])
    # argmax: first maximum
    ties = np.tile(U["argmax_ties"], (3, 1)).astype(np.float32)
    dt = ctx.to_device(ties)
    a3 = ctx.empty((3,), np.uint8)
    chk(lib.dfrl_argmax(ctx.h, dt.p, 3, 8, a3.p))
    assert np.all(a3.get() == int(U["argmax_ties_idx"]))


# ------------------------------------------------------------------------- K4 returns/GAE ---

# (the last two sizes take the 4-envs-per-thread kernel: n % 4 == 0, n >= 16384)
@pytest.mark.parametrize("n,T", [(1, 4), (257, 4), (1000, 8), (64, 33), (16384, 4), (65540, 8)])
def test_gae_vs_oracle(D, ctx, orc, n, T):
    lib, chk = D._lib.lib, D._lib.check
    rng = np.random.default_rng(n + T)
    done = (rng.random((T, n)) < 0.2).astype(np.uint8)
    vs = rng.standard_normal((T, n)).astype(np.float32)
    ve = rng.standard_normal((T, n)).astype(np.float32)
    tg_w, adv_w = orc.gae(done, vs, ve, 0.99, 0.95)
    dd, dvs, dve = ctx.to_device(done), ctx.to_device(vs), ctx.to_device(ve)
    tg, adv = ctx.empty((T, n), np.float32), ctx.empty((T, n), np.float32)
    chk(lib.dfrl_gae(ctx.h, dd.p, dvs.p, dve.p, n, T, 0.99, 0.95, tg.p, adv.p))
    close(tg.get(), tg_w, what="targets")
    close(adv.get(), adv_w, what="advantages")


@pytest.mark.parametrize("n,L", [(3, 66), (300, 40)])
def test_reinforce_returns_vs_oracle(D, ctx, orc, n, L):
    lib, chk = D._lib.lib, D._lib.check
    rng = np.random.default_rng(L)
    done = (rng.random((L, n)) < 0.1).astype(np.uint8)
    length = rng.integers(1, L + 1, n).astype(np.int32)
    for i in range(n):  # every env's record ends with a finished episode
        done[length[i] - 1, i] = 1
    g_w, acc_w = orc.returns(done, length, 0.99)
    dd, dl = ctx.to_device(done), ctx.to_device(length)
    g = ctx.zeros((L, n), np.float32)
    acc = ctx.zeros((2,), np.float64)
    chk(lib.dfrl_returns(ctx.h, dd.p, dl.p, n, L, 0.99, g.p, acc.p))
    mask = np.arange(L)[:, None] < length[None, :]
    close(g.get()[mask], g_w[mask], what="returns")
    a = acc.get()
    assert a[1] == acc_w[1] and abs(a[0] - acc_w[0]) < 1e-6 * abs(acc_w[0])


# ------------------------------------------------------------------------- K5 losses --------

def test_loss_gradients_vs_reference_golden(D, ctx):
    lib, chk = D._lib.lib, D._lib.check
    P, PO, ch, adv = U["loss_p"], U["loss_pold"], U["loss_choice"].astype(np.uint8), U["loss_adv"]
    dP, dPO, dch, dadv = ctx.to_device(P), ctx.to_device(PO), ctx.to_device(ch), ctx.to_device(adv)
    dsel = ctx.to_device(PO[np.arange(40), ch])
    out = ctx.empty((40, 8), np.float32)
    chk(lib.dfrl_loss_grad(ctx.h, 0, dP.p, dch.p, dadv.p, None, 0.0, 40, 8, out.p))
    close(out.get(), U["loss_softmax_log"], what="softmax_log")
    chk(lib.dfrl_loss_grad(ctx.h, 1, dP.p, dch.p, dadv.p, dsel.p, 0.0, 40, 8, out.p))
    close(out.get(), U["loss_clipped"], what="clipped")
    chk(lib.dfrl_loss_grad(ctx.h, 2, dP.p, dch.p, dadv.p, dPO.p, 1.0, 40, 8, out.p))
    close(out.get(), U["loss_kl_beta1"], what="kl")


# ------------------------------------------------------------------------- K7 optimizers ----

@pytest.mark.parametrize("key,kind,wd", [("opt_sgd", 0, 0.0), ("opt_sgd_wd", 0, 1e-3),
                                         ("opt_momentum", 1, 0.0), ("opt_adam", 2, 0.0)])
def test_optimizers_vs_reference_golden(D, ctx, key, kind, wd):
    lib, chk = D._lib.lib, D._lib.check
    p = ctx.to_device(U["opt_p0"])
    state = ctx.zeros((100,), np.float32)
    t = 1.0
    for k, g in enumerate(U["opt_grads"]):
        dg = ctx.to_device(g)
        chk(lib.dfrl_opt_step(ctx.h, kind, p.p, dg.p, state.p, 50, 1e-2, wd, 0.9, 0.999, t))
        t += 1.0
        close(p.get(), U[key][k], what=f"{key} step {k}")


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_optimizers_large_vector_vs_oracle(D, ctx, orc, kind):
    # n >= 65536, n % 4 == 0: the 16-byte vectorised kernel (the golden vectors above take the scalar one)
    lib, chk = D._lib.lib, D._lib.check
    n = (1 << 17) + 8
    rng = np.random.default_rng(kind)
    p0 = rng.standard_normal(n).astype(np.float32)
    want, wstate = p0.copy(), np.zeros(2 * n, np.float32)
    p, state = ctx.to_device(p0), ctx.zeros((2 * n,), np.float32)
    for k in range(3):
        g = rng.standard_normal(n).astype(np.float32)
        dg = ctx.to_device(g)
        chk(lib.dfrl_opt_step(ctx.h, kind, p.p, dg.p, state.p, n, 1e-2, 1e-3 if kind == 0 else 0.0, 0.9, 0.999, 1.0 + k))
        orc.opt_step(kind, want, g, wstate, 1e-2, wd=1e-3 if kind == 0 else 0.0, t=1.0 + k)
        close(p.get(), want, what=f"optimizer {kind} step {k}")
        dg.free()
    p.free(); state.free()


# ------------------------------------------------------------------------- whole learners ---

def _run_case(D, ctx, name, fused):
    c = refcases.load_case(name)
    algo, n, work, iters = int(c["algo"]), int(c["n_envs"]), int(c["work"]), int(c["iters"])
    policy = D.Model(ctx, c["policy_layers"], 32)
    policy.set_parameters(c["pparams0"])
    value = None
    if len(c["value_layers"]):
        value = D.Model(ctx, c["value_layers"], 32)
        value.set_parameters(c["vparams0"])
    env = D.Environment(ctx, n)
    env.set_state(refcases.initial_state(c["steps"], n))
    tr = D.Trainer(ctx, env, policy, value, algo=algo, work=work, policy_lr=float(c["plr"]),
                   value_lr=float(c["vlr"]), policy_opt=int(c["popt"]), value_opt=int(c["vopt"]),
                   policy_wd=float(c["pwd"]), action_mode=D.ACT_FORCED, fused=1 if fused else 0)
    pg_i = vg_i = 0
    epochs = 1 if algo in (D.REINFORCE, D.ACTOR_CRITIC) else 4
    for it in range(iters):
        rec = refcases.records_from_steps(c["steps"], it, n)
        L = rec["L"]
        Lt = tr.read(D.F_REC_ACTION).shape[0]
        items = np.zeros((Lt, n), np.uint8)
        acts = np.zeros((Lt, n), np.uint8)
        items[:L], acts[:L] = rec["items"], rec["action"]
        tr.rollout(items=items, actions=acts)
        mask = np.arange(L)[:, None] < rec["len"][None, :]
        # transitions: bit-exact
        got_state = tr.read(D.F_REC_STATE)[:L]
        assert np.array_equal(got_state.transpose(0, 2, 1)[mask], rec["rec_state"].transpose(0, 2, 1)[mask])
        assert np.array_equal(tr.read(D.F_REC_DONE)[:L][mask], rec["done"][mask])
        assert np.array_equal(tr.read(D.F_REC_ACTION)[:L][mask], rec["action"][mask])
        if algo == D.REINFORCE:
            assert np.array_equal(tr.read(D.F_REC_LEN), rec["len"])
        else:
            assert np.array_equal(env.state(), rec["final_state"])
        # policy outputs at the visited states == the reference's action.distrib
        close(tr.read(D.F_REC_PROBS)[:L][mask], rec["p_old"][mask], what=f"{name} it{it} p_old")
        tr.learn()
        adv_ref = refcases.adv_from_rows(c["rows"], it, n, L)
        close(tr.read(D.F_ADVANTAGE)[:L][mask], adv_ref[mask], what=f"{name} it{it} advantages")
        if value is not None:
            close(tr.read(D.F_VALUE_GRAD), c["value_grads"][vg_i], what=f"{name} it{it} value grad")
            vg_i += 1
        log = tr.read(D.F_POLICY_GRAD_LOG)
        for e in range(epochs):
            try:
                close(log[e], c["policy_grads"][pg_i], what=f"{name} it{it} policy grad {e}")
            except AssertionError:
                # reference initialisation at >= 1024 rows: a relu pre-activation closer to zero than
                # two fp32 summation orders agree on may flip (tests/flipcheck.py); only the fused
                # 3-dense-layer PPO nets are large enough for that to happen
                dims = [int(l[1]) for l in c["policy_layers"] if l[0] == D.DENSE] + [8]
                if algo != D.PPO or len(dims) != 4 or n * L < 1024:
                    raise
                import flipcheck
                from oracle import orc
                before = c["pparams0"] if pg_i == 0 else c["policy_params_log"][pg_i - 1]
                obs = orc.obs_encode(rec["rec_state"].transpose(1, 0, 2).reshape(18, L * n), 8)
                acts = rec["action"].reshape(-1).astype(np.int64)
                advs = adv_ref.reshape(-1).astype(np.float64)
                poa = rec["p_old"].reshape(-1, 8)[np.arange(L * n), acts].astype(np.float64)
                Dm, _ = flipcheck.ambiguous_directions(
                    obs, before, dims, lambda rows, o: flipcheck.policy_dlogits(o, acts[rows], advs[rows], poa[rows], flipcheck.PPO))
                flipcheck.flip_close(log[e], c["policy_grads"][pg_i], Dm, what=f"{name} it{it} policy grad {e}")
            pg_i += 1
        close(policy.parameters(), c["policy_params_log"][pg_i - 1], what=f"{name} it{it} policy params")
        if algo == D.KL_PPO:
            # kl_ppo_learner::beta_ lives on the device (adapted by a one-thread kernel between the steps, read back
            # with the statistics): 1 -> doubled -> clamped to 0.1 after the first step (d_targ = 1e-9,
            # policy_gradient.h:68-83, 333-334); the per-step gradients above already depend on it
            assert tr.stats()["kl_beta"] == pytest.approx(0.1, rel=1e-6)
    close(policy.parameters(), c["pparams_final"], what="final policy params")
    if value is not None:
        close(value.parameters(), c["vparams_final"], what="final value params")
    tr.close(); env.close(); policy.close()
    if value is not None:
        value.close()


@pytest.mark.parametrize("name", refcases.case_names())
def test_trainer_layered_vs_reference_trace(D, ctx, name):
    _run_case(D, ctx, name, fused=False)


@pytest.mark.parametrize("name", refcases.case_names())
def test_trainer_default_path_vs_reference_trace(D, ctx, name):
    # fused kernels where the nets qualify (identical results required), layered otherwise
    _run_case(D, ctx, name, fused=True)


@pytest.mark.parametrize("vend", [0, 1])
@pytest.mark.parametrize("n,T,ctas", [(512, 4, 0), (200, 4, 0), (96, 8, 0), (70, 5, 0),
                                      (512, 4, 3), (544, 4, 2), (200, 4, 1), (416, 8, 4), (2048, 4, 5),
                                      (300, 1, 2), (150, 16, 2), (60, 32, 1), (40, 128, 0)])
def test_trainer_vs_oracle_with_uniform_tape(D, ctx, orc, n, T, ctas, vend):
    # vend = 1: V(end) from the compacted pre-pass kernel (fused_vend_kernel: live units + done-flag scan
    # units; the learner kernels then run without their end pass) -- the path large batches take
    # ctas > 0: the persistent learner kernels run on that many CTAs only, so that every CTA works
    # through several row tiles per tile pipeline (the steady state of the large configurations):
    # (512, 4, 3) 16 tiles on 3 CTAs; (544, 4, 2) 17 tiles: an odd tile count per CTA, the two
    # pipelines get different numbers of tiles; (200, 4, 1) 7 tiles on one CTA with a ragged last
    # tile; (416, 8, 4) 16 envs per tile; (2048, 4, 5) 64 tiles, 12-13 per CTA; (300, 1, 2) single-step
    # rollouts: 128 envs per learner tile, every row is a last-step row; (150, 16, 2) / (60, 32, 1) 8 / 4
    # envs per tile; (40, 128, 0) one env per tile, T = the tile height.
    # sampling mode driven by a tape of uniforms: GPU and oracle must pick identical actions.
    # (200, 4): ragged last tiles of the rollout (128 envs) and learner (32 envs) kernels;
    # (96, 8): 16 envs per learner tile; (70, 5): 25 envs per tile, the byte-wise staging path.
    B = 8
    rng = np.random.default_rng(17)
    pl = D.fc_layers([32, 64, 64, 8], D.SOFTMAX)
    vl = D.fc_layers([32, 64, 64, 1])
    pnet, vnet = orc.Net(pl, 32), orc.Net(vl, 32)
    # relu'(x) is discontinuous at 0: a pre-activation within rounding distance of zero can take
    # either sign under equally valid summation orders (the fused kernels use bf16 hi/lo split
    # operands, ~1e-5 relative) and then moves a whole gradient row. Hidden biases of +-5 with small
    # weights keep every pre-activation far from zero -- half of the units always on, half always
    # off -- so both mask states are exercised and the comparison below is a strict 1e-4 one.
    def safe_params(dims, seed):
        r = np.random.default_rng(seed)
        parts = []
        for li, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
            parts.append((r.standard_normal(a * b) * 0.05).astype(np.float32))
            bias = (r.standard_normal(b) * 0.05).astype(np.float32)
            if li < len(dims) - 2:
                bias += np.where(np.arange(b) % 2 == 0, 5.0, -5.0).astype(np.float32)
            parts.append(bias)
        return np.concatenate(parts)
    pp, vp = safe_params([32, 64, 64, 8], 1), safe_params([32, 64, 64, 1], 2)
    assert pp.size == pnet.param_count() and vp.size == vnet.param_count()
    policy, value = D.Model(ctx, pl, 32), D.Model(ctx, vl, 32)
    policy.set_parameters(pp)
    value.set_parameters(vp)
    first = rng.integers(0, 2, n).astype(np.uint8)
    ecfg = orc.env_cfg(B)
    st = orc.env_reset_all(ecfg, n, first)
    env = D.Environment(ctx, n)
    env.set_state(st)
    # (activations of ~5 make the SUM gradients large: rates scaled down so that 20 updates stay tame)
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=2e-8, value_lr=2e-8,
                   action_mode=D.ACT_SAMPLE)
    if ctas:
        D._lib.check(D._lib.lib.dfrl_debug_set_fused_ctas(tr.h, ctas))
    D._lib.check(D._lib.lib.dfrl_debug_set_vend(tr.h, vend))
    lr = orc.Learner(orc.train_cfg(orc.PPO, T, policy_lr=2e-8, value_lr=2e-8), ecfg, pnet, pp, vnet, vp)
    done_at_last_step = done_mid = 0
    for it in range(5):
        items = rng.integers(0, 2, (T, n)).astype(np.uint8)
        u = rng.random((T, n))
        ro = orc.rollout(ecfg, st, pnet, lr.pparams, T, 0, items, u=u)
        done_at_last_step += int(ro["done"][T - 1].sum())
        done_mid += int(ro["done"][:T - 1].sum())
        tr.rollout(items=items, u=u)
        ga = tr.read(D.F_REC_ACTION)
        agree = np.mean(ga == ro["action"])
        # identical unless a 1e-7 probability difference straddles u (vanishingly rare)
        assert agree == 1.0, agree
        assert np.array_equal(tr.read(D.F_REC_DONE), ro["done"])
        assert np.array_equal(env.state(), st)
        out = lr.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
        tr.learn()
        close(tr.read(D.F_ADVANTAGE), out["adv"], what="adv")
        close(tr.read(D.F_VALUE_TARGET), out["targets"], what="targets")
        close(tr.read(D.F_VALUE_GRAD), out["value_grad"], what="vgrad")
        close(tr.read(D.F_POLICY_GRAD_LOG), out["policy_grads"], what="pgrads")
        close(policy.parameters(), lr.pparams, what="pparams")
        close(value.parameters(), lr.vparams, what="vparams")
    assert done_at_last_step > 0 and (done_mid > 0 or T == 1)  # both kinds of end rows were exercised
    s = tr.stats()
    assert s["env_steps"] == 5 * n * T
    assert s["reward_sum"] + s["episodes"] == s["env_steps"]
    tr.close(); env.close(); policy.close(); value.close()


def test_known_answer_weights20_argmax_eval(D, ctx):
    # deep_agent.cc with weights.20: 26.553 +- 0.028 per 10 000-episode round (deep.log)
    policy = D.Model(ctx, D.conv_layers([4, 128, 64, 1]), 32)
    policy.set_parameters(U["weights20"])
    env = D.Environment(ctx, 8192, seed=2021)
    mean, steps = D.eval_argmax(ctx, env, policy, 4)
    assert 26.45 < mean < 26.65, mean
    assert steps == round(mean * 8192 * 4) + 8192 * 4
    env.close(); policy.close()


def test_free_running_ppo_at_c2_size_invariants(D, ctx):
    # full C2 size, Philox items + sampling, size-independent properties
    n, T = 4096, 4
    policy = D.Model(ctx, D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
    value = D.Model(ctx, D.fc_layers([32, 64, 64, 1]), 32)
    policy.init_parameters(1)
    value.init_parameters(2)
    env = D.Environment(ctx, n, seed=1234)
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-4 * 32 / (n * T),
                   value_lr=1e-5 * 32 / (n * T))
    p0 = policy.parameters()
    tr.iterate(20)
    s = tr.stats()
    assert s["env_steps"] == 20 * n * T
    assert s["reward_sum"] + s["episodes"] == s["env_steps"]
    p1 = policy.parameters()
    assert np.all(np.isfinite(p1)) and np.any(p1 != p0)
    st = env.state()
    assert st[:16].min() >= 0 and st[:16].max() <= 8 and set(np.unique(st[16])) <= {1, 4}
    done = tr.read(D.F_REC_DONE)
    assert 0.02 < done.mean() < 0.2   # near-random policy: ~1 episode end per 12.6 steps
    tr.close(); env.close(); policy.close(); value.close()


def test_free_running_ppo_is_bitwise_reproducible(D, ctx):
    # fixed-order reductions everywhere (per-CTA partials, slice sums): two runs of the same
    # configuration give identical bits, also across the rollout / learner tile boundaries
    def run():
        n, T = 4096 + 40, 4   # not a multiple of the 128-env rollout tile nor of the 32-env learner tile
        policy = D.Model(ctx, D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 64, 1]), 32)
        policy.init_parameters(3)
        value.init_parameters(4)
        env = D.Environment(ctx, n, seed=77)
        tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-4 * 32 / (n * T),
                       value_lr=1e-5 * 32 / (n * T))
        tr.iterate(8)
        out = (policy.parameters().copy(), value.parameters().copy(), env.state().copy(), tr.stats())
        tr.close(); env.close(); policy.close(); value.close()
        return out
    a, b = run(), run()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3]["reward_sum"] == b[3]["reward_sum"] and a[3]["episodes"] == b[3]["episodes"]
    assert np.all(np.isfinite(a[0])) and np.all(np.isfinite(a[1]))


def test_async_stats_pipeline_matches_blocking_reads(D, ctx):
    # stats_begin / stats_end with one step in flight (the e2e loop of bench.py) must return, step
    # for step, what the blocking stats() returns; host tapes are double buffered by the caller
    def run(pipelined):
        n, T, steps = 1000, 4, 6
        rng = np.random.default_rng(5)
        policy = D.Model(ctx, D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
        value = D.Model(ctx, D.fc_layers([32, 64, 64, 1]), 32)
        policy.init_parameters(3)
        value.init_parameters(4)
        env = D.Environment(ctx, n, seed=9)
        tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=1e-6, value_lr=1e-6)
        tapes = [rng.integers(0, 2, (T, n)).astype(np.uint8) for _ in range(steps)]
        out = []
        if pipelined:
            for i in range(steps):
                tr.rollout(items=tapes[i])
                tr.learn()
                tr.stats_begin()
                if i:
                    out.append(tr.stats_end())
            out.append(tr.stats_end())
            with pytest.raises(D._lib.DfrlError):
                tr.stats_end()  # nothing in flight
        else:
            for i in range(steps):
                tr.rollout(items=tapes[i])
                tr.learn()
                out.append(tr.stats())
        params = policy.parameters().copy()
        tr.close(); env.close(); policy.close(); value.close()
        return out, params
    (a, pa), (b, pb) = run(True), run(False)
    assert a == b and a[-1]["env_steps"] == 6 * 1000 * 4
    assert np.array_equal(pa, pb)


def _safe_params(dims, seed, bias=5.0):
    """Hidden biases of +-bias with small weights: every pre-activation is far from zero (half of
    the units always on, half always off), so relu masks cannot flip under rounding differences."""
    r = np.random.default_rng(seed)
    parts = []
    for li, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        parts.append((r.standard_normal(a * b) * 0.05).astype(np.float32))
        bvec = (r.standard_normal(b) * 0.05).astype(np.float32)
        if li < len(dims) - 2:
            bvec += np.where(np.arange(b) % 2 == 0, bias, -bias).astype(np.float32)
        parts.append(bvec)
    return np.concatenate(parts)


def _conv_safe_params(dims, seed, bias=2.0):
    """Mask-stable AND well-conditioned parameters for a conv1d_1 net 4-D1-D2-1. The net's 8 outputs
    per sample go through a softmax, so the output gradients of a sample's 8 rows sum to ~0 and every
    weight gradient depends only on how the activations VARY across a sample's bins: with the +-5
    biases of _safe_params that variation is ~1 % of the activations and rounding differences (2^-17
    relative on the tensor pipe) would be amplified 100 x. Here: hidden biases of +-bias, layer 1
    |W1 . x| <= 4 * 0.4 (x in [0, 1]) < bias, layer 2 weights with zero sum over the always-on units
    of layer 1 (no common-mode shift), so pre-activations stay ~bias +- 0.25 * 6 sigma away from zero
    while the activations vary by ~10-20 % across rows."""
    r = np.random.default_rng(seed)
    d0, d1, d2, d3 = dims
    on1 = np.arange(d1) % 2 == 0
    w1 = r.uniform(-0.4, 0.4, (d1, d0))
    b1 = np.where(on1, bias, -bias) + r.uniform(-0.05, 0.05, d1)
    w2 = r.uniform(-0.1, 0.1, (d2, d1))
    w2[:, on1] -= w2[:, on1].mean(axis=1, keepdims=True)
    b2 = np.where(np.arange(d2) % 2 == 0, bias, -bias) + r.uniform(-0.05, 0.05, d2)
    w3 = r.uniform(-0.3, 0.3, (d3, d2))
    b3 = r.uniform(-0.05, 0.05, d3)
    return np.concatenate([w1.ravel(), b1, w2.ravel(), b2, w3.ravel(), b3]).astype(np.float32)


@pytest.mark.parametrize("algo_name,n,T,B,pdims,vdims", [
    # 32 bins (O = 128): layered path; 1200 x 4 = 4800 learner rows put the 64-wide products on
    # the tcgen05 GEMMs of gemm_umma.cu, the rest on the FFMA kernels
    ("ppo", 1200, 4, 32, [128, 64, 64, 32], [128, 64, 64, 1]),
    # online actor-critic on the fused kernels (softmax-CE head: identity backward), T = 8
    ("ac", 300, 8, 8, [32, 64, 64, 8], [32, 64, 64, 1]),
    # the same on 2 CTAs: 69 learner tiles of 16 envs, 9 rollout tiles -> every tile pipeline of the
    # rollout (4), GAE (4), critic (2) and policy (2) kernels works through several tiles
    ("ac@2", 1100, 8, 8, [32, 64, 64, 8], [32, 64, 64, 1]),
    # 16-wide hidden layers (the second instantiation of the fused kernels: one epilogue thread per
    # row in the critic step, K = 16 single-chunk TMEM operands), PPO, 3 CTAs
    ("ppo@3", 700, 4, 8, [32, 16, 16, 8], [32, 16, 16, 1]),
    # the reference's own critic 32-64-32-1 (ppo_training.cc:19-26) on the fused critic / GAE / V(end) kernels
    ("ppo@3", 650, 4, 8, [32, 64, 64, 8], [32, 64, 32, 1]),
    # (the reference's conv1d policies on the fused kernels: tests/test_gpu_scale.py, flip-aware, He init --
    #  mask-stable "safe" parameters make a conv net's gradient ill-conditioned: see _conv_safe_params)
])
def test_trainer_variants_vs_oracle(D, ctx, orc, algo_name, n, T, B, pdims, vdims):
    algo_name, _, cap = algo_name.partition("@")
    algo, oalgo = (D.PPO, orc.PPO) if algo_name == "ppo" else (D.ACTOR_CRITIC, orc.ACTOR_CRITIC)
    last = D.SOFTMAX if algo_name == "ppo" else D.SOFTMAX_CE
    rng = np.random.default_rng(5)
    conv = isinstance(pdims, tuple)
    if conv:
        pdims = pdims[1]
    pl, vl = (D.conv_layers if conv else D.fc_layers)(pdims, last), D.fc_layers(vdims)
    pnet, vnet = orc.Net(pl, 4 * B), orc.Net(vl, 4 * B)
    pp, vp = (_conv_safe_params(pdims, 11) if conv else _safe_params(pdims, 11)), _safe_params(vdims, 12)
    policy, value = D.Model(ctx, pl, 4 * B), D.Model(ctx, vl, 4 * B)
    policy.set_parameters(pp)
    value.set_parameters(vp)
    ecfg = orc.env_cfg(B)
    st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
    env = D.Environment(ctx, n, n_bins=B)
    env.set_state(st)
    tr = D.Trainer(ctx, env, policy, value, algo=algo, work=T, policy_lr=2e-8, value_lr=2e-8,
                   action_mode=D.ACT_SAMPLE)
    if conv:  # the whole iteration runs on the fused kernels (fused_conv.cuh + the fused critic kernels)
        assert tr.fused_covers_iteration()
    if cap:
        D._lib.check(D._lib.lib.dfrl_debug_set_fused_ctas(tr.h, int(cap)))
        D._lib.check(D._lib.lib.dfrl_debug_set_vend(tr.h, 1))  # the capped cases also take the compacted V(end) pre-pass
    lr = orc.Learner(orc.train_cfg(oalgo, T, policy_lr=2e-8, value_lr=2e-8), ecfg, pnet, pp, vnet, vp)
    for it in range(3):
        items = rng.integers(0, 2, (T, n)).astype(np.uint8)
        u = rng.random((T, n))
        ro = orc.rollout(ecfg, st, pnet, lr.pparams, T, 0, items, u=u)
        tr.rollout(items=items, u=u)
        assert np.array_equal(tr.read(D.F_REC_ACTION), ro["action"])
        assert np.array_equal(tr.read(D.F_REC_DONE), ro["done"])
        assert np.array_equal(env.state(), st)
        close(tr.read(D.F_REC_PROBS), ro["probs"], what="p_old")
        out = lr.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
        tr.learn()
        close(tr.read(D.F_ADVANTAGE), out["adv"], what="adv")
        close(tr.read(D.F_VALUE_GRAD), out["value_grad"], what="vgrad")
        close(tr.read(D.F_POLICY_GRAD_LOG), out["policy_grads"], what="pgrads")
        close(policy.parameters(), lr.pparams, what="pparams")
        close(value.parameters(), lr.vparams, what="vparams")
    tr.close(); env.close(); policy.close(); value.close()


def test_abi_error_paths(D, ctx):
    """Status codes + dfrl_last_error(), the convention the C++ mirror rethrows as xeno::error
    (reference: shape checks throwing xeno::error, tensor.cc:41-69)."""
    import ctypes as C
    lib = D._lib.lib
    h = C.c_void_p()
    # model whose widths do not chain
    kinds = (C.c_int * 2)(D.DENSE, D.DENSE)
    ins, outs = (C.c_int * 2)(32, 48), (C.c_int * 2)(64, 8)
    rc = lib.dfrl_mlp_create(ctx.h, 2, kinds, ins, outs, 32, C.byref(h))
    assert rc == -1 and lib.dfrl_last_error()
    # environment id / action out of range (per-id triple)
    env = D.Environment(ctx, 4)
    assert lib.dfrl_env_apply_one(env.h, 7, 0) == -1 and b"out of range" in lib.dfrl_last_error()
    assert lib.dfrl_env_apply_one(env.h, 0, 9) == -1
    # trainer: policy output width != bins
    bad = D.Model(ctx, D.fc_layers([32, 16, 4], D.SOFTMAX), 32)
    val = D.Model(ctx, D.fc_layers([32, 16, 1]), 32)
    with pytest.raises(D._lib.DfrlError, match="action.cardinality"):
        D.Trainer(ctx, env, bad, val, algo=D.PPO, work=4)
    # forced mode without an action tape, learn before rollout is allowed to run on empty records
    good = D.Model(ctx, D.fc_layers([32, 16, 8], D.SOFTMAX), 32)
    tr = D.Trainer(ctx, env, good, val, algo=D.PPO, work=4, action_mode=D.ACT_FORCED)
    with pytest.raises(D._lib.DfrlError, match="action tape"):
        tr.rollout()
    with pytest.raises(D._lib.DfrlError, match="teacher-force"):
        tr.iterate(1)
    tr.close(); env.close(); bad.close(); val.close(); good.close()


def test_single_environment_trainer(D, ctx, orc):
    # N = 1 (BASELINE configs[0] shape): one environment, every kernel with a single row tile
    pl, vl = D.fc_layers([32, 64, 64, 8], D.SOFTMAX), D.fc_layers([32, 64, 64, 1])
    pnet, vnet = orc.Net(pl, 32), orc.Net(vl, 32)
    pp, vp = _safe_params([32, 64, 64, 8], 1), _safe_params([32, 64, 64, 1], 2)
    policy, value = D.Model(ctx, pl, 32), D.Model(ctx, vl, 32)
    policy.set_parameters(pp)
    value.set_parameters(vp)
    ecfg = orc.env_cfg(8)
    st = orc.env_reset_all(ecfg, 1, np.array([1], np.uint8))
    env = D.Environment(ctx, 1)
    env.set_state(st)
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=4, policy_lr=1e-7, value_lr=1e-7)
    lr = orc.Learner(orc.train_cfg(orc.PPO, 4, policy_lr=1e-7, value_lr=1e-7), ecfg, pnet, pp, vnet, vp)
    rng = np.random.default_rng(9)
    for it in range(6):
        items = rng.integers(0, 2, (4, 1)).astype(np.uint8)
        u = rng.random((4, 1))
        ro = orc.rollout(ecfg, st, pnet, lr.pparams, 4, 0, items, u=u)
        tr.rollout(items=items, u=u)
        assert np.array_equal(tr.read(D.F_REC_ACTION), ro["action"])
        assert np.array_equal(env.state(), st)
        out = lr.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
        tr.learn()
        close(tr.read(D.F_POLICY_GRAD_LOG), out["policy_grads"], what="pgrads")
        close(value.parameters(), lr.vparams, what="vparams")
    tr.close(); env.close(); policy.close(); value.close()
