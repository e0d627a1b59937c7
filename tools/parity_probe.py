"""Exploration tool (not a test): raw norm-wise / elementwise errors of the fused path against the
double-accumulating oracle at the BASELINE sizes with the reference's initialisation.
  python tools/parity_probe.py N T ITERS [ppo|ac]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dependence_free_rl_b200 as D  # noqa: E402
from oracle import orc  # noqa: E402

n, T, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
algo = sys.argv[4] if len(sys.argv) > 4 else "ppo"
B = 8
ctx = D.Context(0)
last = D.SOFTMAX if algo == "ppo" else D.SOFTMAX_CE
pl, vl = D.fc_layers([32, 64, 64, 8], last), D.fc_layers([32, 64, 64, 1])
pnet, vnet = orc.Net(pl, 32), orc.Net(vl, 32)
policy, value = D.Model(ctx, pl, 32), D.Model(ctx, vl, 32)
policy.init_parameters(1)
value.init_parameters(2)
pp, vp = policy.parameters(), value.parameters()
rng = np.random.default_rng(3)
ecfg = orc.env_cfg(B)
st = orc.env_reset_all(ecfg, n, rng.integers(0, 2, n).astype(np.uint8))
for _ in range(6):  # spread the envs over their episodes
    orc.env_step(ecfg, st, rng.integers(0, B, n).astype(np.uint8), rng.integers(0, 2, n).astype(np.uint8))
env = D.Environment(ctx, n)
env.set_state(st)
plr, vlr = 1e-4 * 32 / (n * T), 1e-5 * 32 / (n * T)
dalgo, oalgo = (D.PPO, orc.PPO) if algo == "ppo" else (D.ACTOR_CRITIC, orc.ACTOR_CRITIC)
tr = D.Trainer(ctx, env, policy, value, algo=dalgo, work=T, policy_lr=plr, value_lr=vlr)
lr = orc.Learner(orc.train_cfg(oalgo, T, policy_lr=plr, value_lr=vlr), ecfg, pnet, pp, vnet, vp, f64="mt")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300), np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


for it in range(iters):
    items = rng.integers(0, 2, (T, n)).astype(np.uint8)
    u = rng.random((T, n))
    t0 = time.time()
    ro = orc.rollout(ecfg, st, pnet, lr.pparams, T, 0, items, u=u)
    t1 = time.time()
    tr.rollout(items=items, u=u)
    ga, gd = tr.read(D.F_REC_ACTION), tr.read(D.F_REC_DONE)
    print(f"it {it}: actions differ {int((ga != ro['action']).sum())} of {ga.size}, done differ "
          f"{int((gd != ro['done']).sum())}, state equal {np.array_equal(env.state(), st)}; p_old",
          rel(tr.read(D.F_REC_PROBS), ro["probs"]))
    if not np.array_equal(ga, ro["action"]):
        print("  (u-straddle: stopping the bit-exact comparison here)")
        break
    out = lr.learn(ro["state"], st, ro["action"], ro["done"], ro["probs"])
    t2 = time.time()
    tr.learn()
    print(f"  oracle rollout {t1 - t0:.1f}s learn {t2 - t1:.1f}s")
    print("  adv", rel(tr.read(D.F_ADVANTAGE), out["adv"]), "targets", rel(tr.read(D.F_VALUE_TARGET), out["targets"]))
    print("  vgrad", rel(tr.read(D.F_VALUE_GRAD), out["value_grad"]))
    log = tr.read(D.F_POLICY_GRAD_LOG)
    for e in range(log.shape[0]):
        print("  pgrad", e, rel(log[e], out["policy_grads"][e]))
    print("  pparams", rel(policy.parameters(), lr.pparams), "vparams", rel(value.parameters(), lr.vparams))
    print("  update size: policy", rel(policy.parameters(), pp)[0], "value", rel(value.parameters(), vp)[0])
