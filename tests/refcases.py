"""Shared helpers for the parity tests: reference-trace -> dfrl record layouts, tolerances,
and the list of golden cases generated from the compiled reference (oracle/_ref).

Layouts follow include/dfrl.h: state int8 [2B+2][N]; records step-major [L][N].
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
NB = 8

# fp32 tolerance stated by BASELINE.json north_star: 1e-4 relative.
RTOL = 1e-4


def close(a, b, rtol=RTOL, what=""):
    """|a-b| <= rtol * (|b| + scale) elementwise, scale = max|b| (so tiny entries of a vector
    are judged against the vector's magnitude), plus a norm-wise check."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.all(np.isfinite(a)), f"{what}: non-finite values"
    scale = np.max(np.abs(b)) if b.size else 0.0
    err = np.abs(a - b)
    bound = rtol * (np.abs(b) + scale)
    if not np.all(err <= bound + 1e-30):
        i = np.argmax(err - bound)
        raise AssertionError(f"{what}: max violation at {i}: got {a.flat[i]!r} want {b.flat[i]!r} "
                             f"(scale {scale:.3e}, rel-to-scale {err.flat[i] / (scale + 1e-30):.3e})")
    nb = np.linalg.norm(b)
    if nb > 0:
        rel = np.linalg.norm(a - b) / nb
        assert rel <= rtol, f"{what}: norm-wise relative error {rel:.3e} > {rtol}"


def close_bulk(a, b, rtol=RTOL, max_outlier_frac=2e-3, what=""):
    """For large random problems: a relu pre-activation within one ulp of zero can take a
    different sign under a different (equally valid) summation order, which moves the few
    gradient entries fed by that one row by more than 1e-4.  Require the norm-wise error and
    all but a 2e-3 fraction of the entries (one flipped unit touches one row of dW, ~1e-3 of a layer) to meet the tolerance."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape and np.all(np.isfinite(a)), what
    scale = np.max(np.abs(b))
    bad = np.abs(a - b) > rtol * (np.abs(b) + scale)
    assert bad.mean() <= max_outlier_frac, f"{what}: {bad.mean():.2e} of entries out of tolerance"
    rel = np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel <= rtol, f"{what}: norm-wise relative error {rel:.3e}"


def item_code(item_wh):
    """(4,2) -> 1 (shape1), (1,2) -> 0 (shape2) (bin_packing.h:73-79)."""
    return (np.asarray(item_wh)[..., 0] == 4).astype(np.uint8)


def planes(bins, item):
    """bins [..., B, 2], item [..., 2] -> int8 [2B+2, ...] plane layout."""
    bins = np.asarray(bins)
    item = np.asarray(item)
    lead = bins.shape[:-2]
    B = bins.shape[-2]
    out = np.zeros((2 * B + 2,) + lead, dtype=np.int8)
    for b in range(B):
        out[2 * b] = bins[..., b, 0]
        out[2 * b + 1] = bins[..., b, 1]
    out[2 * B] = item[..., 0]
    out[2 * B + 1] = item[..., 1]
    return out


def records_from_steps(steps, it, n_envs, B=NB, cap=(8, 8)):
    """Reference step log of iteration `it` -> dict of dfrl-layout records.

    Works for fixed-length (AC/PPO) and variable-length (REINFORCE) rollouts: L = max steps of
    any env, shorter envs are zero padded and `len` says how many are valid."""
    s = steps[steps["iter"] == it]
    length = np.zeros(n_envs, dtype=np.int32)
    for e in range(n_envs):
        length[e] = int(np.sum(s["env"] == e))
    L = int(length.max())
    P = 2 * B + 2
    rec_state = np.zeros((L, P, n_envs), np.int8)
    end_state = np.zeros((L, P, n_envs), np.int8)
    action = np.zeros((L, n_envs), np.uint8)
    done = np.zeros((L, n_envs), np.uint8)
    items = np.zeros((L, n_envs), np.uint8)
    p_old = np.zeros((L, n_envs, B), np.float32)
    final_state = np.zeros((P, n_envs), np.int8)
    for r in s:
        e, t = int(r["env"]), int(r["t"])
        rec_state[t, :, e] = planes(r["sbins"], r["sitem"])
        end_state[t, :, e] = planes(r["ebins"], r["eitem"])
        action[t, e] = r["action"]
        done[t, e] = r["done"]
        items[t, e] = item_code(r["item_after"])
        p_old[t, e] = r["p_old"]
        if t == length[e] - 1:
            if r["done"]:
                fresh = np.tile(np.array(cap, np.int32), (B, 1))
                final_state[:, e] = planes(fresh, r["item_after"])
            else:
                final_state[:, e] = end_state[t, :, e]
    return {"rec_state": rec_state, "end_state": end_state, "action": action, "done": done,
            "items": items, "p_old": p_old, "final_state": final_state, "len": length, "L": L}


def initial_state(steps, n_envs, B=NB, cap=(8, 8)):
    s = steps[steps["iter"] == -1]
    st = np.zeros((2 * B + 2, n_envs), np.int8)
    for r in s:
        st[:, int(r["env"])] = planes(r["sbins"], r["sitem"])
    return st


def adv_from_rows(rows, it, n_envs, L):
    """Reference per-row advantages of iteration `it` -> [L][N] (end rows dropped)."""
    r = rows[(rows["iter"] == it) & (rows["t"] >= 0)]
    adv = np.zeros((L, n_envs), np.float32)
    adv[r["t"], r["env"]] = r["advantage"]
    return adv


# ------------------------------------------------------------------ golden case catalogue ----
# (name, algo, seed, n_envs, work, iters, policy layers, value layers, plr, vlr, popt, vopt, pwd)
def _cases():
    from oracle import ref as R
    fc, conv = R.fc_net, R.conv_net
    return [
        dict(name="ppo_small", algo=R.PPO, seed=11, n_envs=6, work=4, iters=3,
             policy=fc([32, 16, 16, 8], R.SOFTMAX), value=fc([32, 16, 8, 1]),
             plr=1e-3, vlr=1e-3),
        dict(name="ppo_c2net", algo=R.PPO, seed=1234, n_envs=8, work=4, iters=1,
             policy=fc([32, 64, 64, 8], R.SOFTMAX), value=fc([32, 64, 64, 1]),
             plr=1e-4, vlr=1e-5),
        dict(name="ppo_refnet_conv", algo=R.PPO, seed=5, n_envs=4, work=4, iters=1,
             policy=conv([4, 128, 64, 1], R.SOFTMAX), value=fc([32, 64, 32, 1]),
             plr=1e-4, vlr=1e-5),
        dict(name="ac_conv_small", algo=R.ACTOR_CRITIC, seed=21, n_envs=5, work=8, iters=3,
             policy=conv([4, 16, 8, 1], R.SOFTMAX_CE), value=fc([32, 16, 8, 1]),
             plr=1e-3, vlr=1e-3),
        dict(name="reinforce_small", algo=R.REINFORCE, seed=31, n_envs=3, work=2, iters=3,
             policy=fc([32, 24, 12, 8], R.SOFTMAX_CE), value=None, plr=1e-4, vlr=0.0),
        dict(name="klppo_small", algo=R.KL_PPO, seed=41, n_envs=4, work=4, iters=2,
             policy=conv([4, 16, 8, 1], R.SOFTMAX), value=fc([32, 16, 8, 1]),
             plr=1e-3, vlr=1e-3, pwd=1e-5),
        dict(name="ppo_adam_small", algo=R.PPO, seed=51, n_envs=4, work=4, iters=3,
             policy=fc([32, 16, 8], R.SOFTMAX), value=fc([32, 16, 1]),
             plr=1e-3, vlr=1e-3, popt=R.ADAM, vopt=R.MOMENTUM),
        # the fused C2 / C4 nets at a size where every fused kernel runs several row tiles, three
        # iterations with the reference's initialisation and its rates scaled to SUMS over 1024 rows
        dict(name="ppo_c2net_256", algo=R.PPO, seed=71, n_envs=256, work=4, iters=3,
             policy=fc([32, 64, 64, 8], R.SOFTMAX), value=fc([32, 64, 64, 1]),
             plr=1e-4 * 32 / 1024, vlr=1e-5 * 32 / 1024),
        # long enough for several episode ends inside rollouts (random-ish policy: ~12 steps)
        dict(name="ppo_long", algo=R.PPO, seed=61, n_envs=4, work=16, iters=2,
             policy=fc([32, 8, 8], R.SOFTMAX), value=fc([32, 8, 1]), plr=1e-4, vlr=1e-4),
    ]


def generate_case(c):
    """Runs the reference (needs oracle/_ref) and returns the arrays of one golden fixture."""
    from oracle import ref as R
    pp = R.init_params(c["policy"], c["seed"] + 1000)
    vp = R.init_params(c["value"], c["seed"] + 2000) if c["value"] is not None else None
    res = R.train(c["algo"], c["seed"], c["n_envs"], c["work"], c["iters"], c["policy"], pp,
                  c["plr"], c["value"], vp, c["vlr"], popt=c.get("popt", 0), vopt=c.get("vopt", 0),
                  pwd=c.get("pwd", 0.0), gamma=0.99, record=True)
    out = {
        "algo": np.int32(c["algo"]), "seed": np.int32(c["seed"]), "n_envs": np.int32(c["n_envs"]),
        "work": np.int32(c["work"]), "iters": np.int32(c["iters"]),
        "policy_layers": np.array(c["policy"].layers, np.int32),
        "value_layers": np.array(c["value"].layers if c["value"] is not None else np.zeros((0, 3)), np.int32),
        "plr": np.float32(c["plr"]), "vlr": np.float32(c["vlr"]),
        "popt": np.int32(c.get("popt", 0)), "vopt": np.int32(c.get("vopt", 0)),
        "pwd": np.float32(c.get("pwd", 0.0)),
        "pparams0": pp, "vparams0": vp if vp is not None else np.zeros(0, np.float32),
        "steps": res["steps"], "rows": res["rows"],
        "opt_iter": np.array([o["iter"] for o in res["opt_log"]], np.int32),
        "opt_which": np.array([o["which"] for o in res["opt_log"]], np.int32),
        "pparams_final": res["policy_params"], "vparams_final": res["value_params"],
    }
    pg = [o["grad"] for o in res["opt_log"] if o["which"] == 0]
    vg = [o["grad"] for o in res["opt_log"] if o["which"] == 1]
    out["policy_grads"] = np.stack(pg) if pg else np.zeros((0, 0), np.float32)
    out["value_grads"] = np.stack(vg) if vg else np.zeros((0, 0), np.float32)
    out["policy_params_log"] = np.stack([o["params"] for o in res["opt_log"] if o["which"] == 0])
    return out


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"train_{name}.npz")))


def case_names():
    return ["ppo_small", "ppo_c2net", "ppo_refnet_conv", "ac_conv_small", "reinforce_small",
            "klppo_small", "ppo_adam_small", "ppo_long", "ppo_c2net_256"]


def orc_net_from_layers(layers, B=NB):
    from oracle import orc
    layers = [tuple(int(v) for v in l) for l in layers]
    first = layers[0]
    input_cols = first[1] if first[0] == orc.DENSE else first[1] * B
    return orc.Net(layers, input_cols)
