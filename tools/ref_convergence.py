"""Learning curve of the UNMODIFIED reference PPO trainer (oracle/_ref, dev container only): the
reference's own nets, rates and schedule (ppo_training.cc:9-31: conv1d 4-128-64-1 softmax policy,
FC 32-64-32-1 critic, SGD 1e-4 / 1e-5, 8 workers x 4 steps per round), evaluated like its main does
(argmax policy, ppo_training.cc:67-81) every CHUNK rounds. Writes profiles/r02_convergence_ref_cpu.csv.
Test infrastructure / evidence only; the product never runs this.
    python tools/ref_convergence.py [total_rounds] [chunk]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 250


def he_fc(dims, seed):
    """FC nets with He-scaled weights (what the fused-net curve of tools/convergence.py uses: the
    reference's N(0, 0.01) dense init, nn.h:12-14, leaves a 3-layer FC net without signal for ~10^5
    rounds)."""
    r = np.random.default_rng(seed)
    return np.concatenate([np.concatenate([(r.standard_normal(a * b) * np.sqrt(2.0 / a)).astype(np.float32),
                                           np.zeros(b, np.float32)]) for a, b in zip(dims[:-1], dims[1:])])


out = os.path.join(ROOT, "profiles", "r02_convergence_ref_cpu.csv")
with open(out, "w") as f:
    f.write("nets,rounds,env_steps,mean_reward_argmax_200ep,seconds\n")
    for name, pol, val, pp, vp in (
            ("reference", R.conv_net([4, 128, 64, 1], R.SOFTMAX), R.fc_net([32, 64, 32, 1]), None, None),
            ("c2_he_init", R.fc_net([32, 64, 64, 8], R.SOFTMAX), R.fc_net([32, 64, 64, 1]),
             he_fc([32, 64, 64, 8], 1234), he_fc([32, 64, 64, 1], 1235))):
        if pp is None:
            pp, vp = R.init_params(pol, 1234), R.init_params(val, 1235)
        t0 = time.time()
        done = 0
        mean, _ = R.eval_argmax(99, pol, pp, 200)
        f.write(f"{name},0,0,{mean:.4f},0.0\n")
        f.flush()
        while done < total:
            res = R.train(R.PPO, 1000 + done, 8, 4, chunk, pol, pp, 1e-4, val, vp, 1e-5, record=False, threads=8)
            pp, vp = res["policy_params"], res["value_params"]
            done += chunk
            mean, _ = R.eval_argmax(99, pol, pp, 200)
            f.write(f"{name},{done},{done * 32},{mean:.4f},{time.time() - t0:.1f}\n")
            f.flush()
