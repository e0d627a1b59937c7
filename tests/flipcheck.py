"""Flip-aware gradient comparison for large batches with the reference's initialisation.

relu'(x) is discontinuous at 0 (nn.h:366-376 masks by pre-activation > 0). With the reference's
N(0, 0.01) weights and zero biases the hidden pre-activations are centred on zero, so among
~10^8 (row, unit) pairs of a 131 072-env batch a few hundred lie closer to zero than ANY two fp32
evaluation orders agree on (the reference itself is -ffast-math + AVX partial sums; the fused
kernels use bf16 hi/lo split products, ~2^-18 relative). Such a unit may legitimately come out
"on" in one implementation and "off" in the other, and since its gradient contribution is a whole
row of dW it shifts a noise-like policy gradient by more than 1e-4 of its norm. Nothing else is
affected: the forward value changes by at most the rounding error itself.

This module makes that statement checkable instead of loosening the tolerance:

  * ambiguous units = (row, layer, unit) with |z| <= KAPPA * E, E = sqrt(sum_k (a_k w_k)^2) the
    natural rounding scale of the dot product (KAPPA = 2^-16, about 6 sigma of the split-product
    rounding error);
  * for every ambiguous unit the exact change of the flat gradient if its mask flipped is computed
    (one row's backward pass) -> direction matrix D [P x m], m << P;
  * the implementation under test must satisfy  g_test = g_oracle + D s + r  with
    |r| <= 1e-4 (|g| + max|g|) elementwise and ||r|| <= 1e-4 ||g||, s from least squares.

m is a few hundred to ~1500 against P = 6792 parameters, so an error outside the span of the
ambiguous rows' own gradient contributions cannot be absorbed.

Nets: dense D0 - relu - D1 - relu - D2 (the C2 / C4 fused nets), flat parameters in the reference
order [W out x in][b out] per layer (nn.h:56-59).
"""
import numpy as np

KAPPA = 2.0 ** -16
PPO, AC = "ppo", "ac"


def split_params(p, dims):
    out, off = [], 0
    for a, b in zip(dims[:-1], dims[1:]):
        W = np.asarray(p[off:off + a * b], np.float64).reshape(b, a)
        off += a * b
        bias = np.asarray(p[off:off + b], np.float64)
        off += b
        out.append((W, bias))
    assert off == len(p)
    return out


def _offsets(dims):
    offs, off = [], 0
    for a, b in zip(dims[:-1], dims[1:]):
        offs.append((off, off + a * b))
        off += (a + 1) * b
    return offs, off


def policy_dlogits(out, actions, adv, p_old_a, kind):
    """Loss gradient at the logits of the selected rows (float64 restatement of rl.h:45-74 +
    nn.h:393-417 for the softmax head / identity for softmax-CE)."""
    p = np.exp(out)
    p /= p.sum(1, keepdims=True)
    idx = np.arange(len(actions))
    if kind == PPO:
        pa = p[idx, actions]
        ratio = pa / p_old_a
        clipped = np.clip(ratio, 0.8, 1.2)
        g = -np.minimum(clipped * adv, ratio * adv) / pa
        dl = -p * (pa * g)[:, None]            # p_j (0 - p_a g_a) for j != a
        dl[idx, actions] += pa * g             # + p_a g_a for j == a
        return dl
    dl = p * adv[:, None]
    dl[idx, actions] -= adv
    return dl


def ambiguous_directions(obs, params, dims, dout_fn, kappa=KAPPA, chunk=1 << 16):
    """obs [R, D0] float32; dout_fn(rows_idx, out[rows]) -> dOut [len, D3] (gradient at the net
    output of those rows). Returns D [P, m] float64 and the list of (row, layer, unit)."""
    (W1, b1), (W2, b2), (W3, b3) = split_params(params, dims)
    R = obs.shape[0]
    cand = []
    for c0 in range(0, R, chunk):
        x = np.asarray(obs[c0:c0 + chunk], np.float64)
        z1 = x @ W1.T + b1
        e1 = np.sqrt((x * x) @ (W1 * W1).T)
        h1 = np.maximum(z1, 0.0)
        z2 = h1 @ W2.T + b2
        e2 = np.sqrt((h1 * h1) @ (W2 * W2).T)
        # (a unit whose inputs are all exactly zero has z = b exactly: never ambiguous)
        for layer, z, e in ((1, z1, e1), (2, z2, e2)):
            r, u = np.nonzero((np.abs(z) <= kappa * e) & (e > 0))
            cand += [(int(c0 + ri), layer, int(ui)) for ri, ui in zip(r, u)]
    offs, P = _offsets(dims)
    if not cand:
        return np.zeros((P, 0)), cand
    rows = np.array(sorted({c[0] for c in cand}))
    pos = {int(r): i for i, r in enumerate(rows)}
    x = np.asarray(obs[rows], np.float64)
    z1 = x @ W1.T + b1
    m1 = z1 > 0
    h1 = np.maximum(z1, 0.0)
    z2 = h1 @ W2.T + b2
    m2 = z2 > 0
    h2 = np.maximum(z2, 0.0)
    out = h2 @ W3.T + b3
    dout = np.asarray(dout_fn(rows, out), np.float64)
    dh2 = dout @ W3
    dh1 = (dh2 * m2) @ W2
    D = np.zeros((P, len(cand)))
    (w1a, w1b), (w2a, w2b), _ = offs
    d0, d1, d2 = dims[0], dims[1], dims[2]
    for j, (r, layer, u) in enumerate(cand):
        i = pos[r]
        if layer == 2:
            a = dh2[i, u] * (-1.0 if m2[i, u] else 1.0)
            D[w2a + u * d1:w2a + (u + 1) * d1, j] = a * h1[i]
            D[w2b + u, j] = a
            dz1 = a * W2[u] * m1[i]
            D[w1a:w1b, j] = np.outer(dz1, x[i]).ravel()
            D[w1b:w1b + d1, j] = dz1
        else:
            a = dh1[i, u] * (-1.0 if m1[i, u] else 1.0)
            D[w1a + u * d0:w1a + (u + 1) * d0, j] = a * x[i]
            D[w1b + u, j] = a
    return D, cand


def flip_close(got, want, D, rtol=1e-4, what=""):
    """got = want + D s + r with r inside the 1e-4 tolerance (elementwise against the vector's
    magnitude and norm-wise). Returns a small report."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape and np.all(np.isfinite(got)), what
    r = got - want
    raw = np.linalg.norm(r) / np.linalg.norm(want)
    used = 0
    if D.shape[1]:
        s, *_ = np.linalg.lstsq(D, r, rcond=None)
        r = r - D @ s
        used = int(np.sum(np.abs(s) > 0.5))
    scale = np.max(np.abs(want))
    rel = np.linalg.norm(r) / np.linalg.norm(want)
    bad = np.abs(r) > rtol * (np.abs(want) + scale)
    assert rel <= rtol, (f"{what}: norm-wise error {rel:.3e} after removing {D.shape[1]} relu-ambiguous directions "
                         f"(raw {raw:.3e})")
    assert not bad.any(), f"{what}: {int(bad.sum())} entries out of tolerance after the projection (raw {raw:.3e})"
    return {"raw": raw, "residual": rel, "ambiguous": D.shape[1], "flipped": used}


def nudge_uniforms(rollout_fn, u, probs_of, rng, margin=1e-5, max_rounds=6):
    """The sampling tape u [T][n] must not sit within `margin` of a CDF boundary of the policy
    output at the state it is used in (there a 1e-7 difference in a probability picks another
    action: the 'u-straddle'). Environments with a close call get fresh uniforms; their
    trajectories are independent of every other environment's. rollout_fn(u) -> rollout dict."""
    for _ in range(max_rounds):
        ro = rollout_fn(u)
        p = np.asarray(probs_of(ro), np.float64)                       # [T][n][B]
        cdf = np.cumsum(p / p.sum(-1, keepdims=True), -1)[..., :-1]    # libstdc++: last entry forced to 1
        near = (np.abs(cdf - u[..., None]) < margin).any(-1).any(0)   # [n]
        if not near.any():
            return ro, u
        u = u.copy()
        u[:, near] = rng.random((u.shape[0], int(near.sum())))
    raise AssertionError("could not find a sampling tape without close calls")
