#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled as oracle/_ref.

Run in the dev container only (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these fixtures --
outputs of the reference's own classes on seeded inputs -- are what pins oracle/dfrl_oracle.c
and, through it, the CUDA path.  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref as R  # noqa: E402
import refcases  # noqa: E402


def units():
    rng = np.random.default_rng(7)
    out = {}
    # engine + distributions (tensor.cc:71-75, 467-470; bin_packing.h:81)
    out["engine_seed1"] = R.engine_draw(1, 16)
    out["engine_seed1234"] = R.engine_draw(1234, 16)
    w = np.array([0.05, 0.3, 0.1, 0.2, 0.05, 0.1, 0.15, 0.05], np.float32)
    out["disc_w"] = w
    out["disc_samples_seed9"] = R.discrete_sample(9, w, 256)
    w2 = rng.random(8).astype(np.float32)
    out["disc_w2"] = w2
    out["disc_samples2_seed77"] = R.discrete_sample(77, w2, 256)
    ties = np.array([0.1, 0.7, 0.7, 0.2, 0.7, 0.0, -1.0, 0.3], np.float32)
    out["argmax_ties"] = ties
    out["argmax_ties_idx"] = np.int32(R.argmax(ties))

    # layers (nn.h:60-110, 113-194, 350-431)
    for name, kind, n_in, n_out, rows, xcols, ycols in [
        ("dense_32_64", R.DENSE, 32, 64, 37, 32, 64),
        ("dense_64_1", R.DENSE, 64, 1, 19, 64, 1),
        ("dense_5_3", R.DENSE, 5, 3, 4, 5, 3),
        ("conv_4_16", R.CONV1D, 4, 16, 9, 32, 128),
        ("conv_16_1", R.CONV1D, 16, 1, 9, 128, 8),
    ]:
        p = (rng.standard_normal((n_in + 1) * n_out) * 0.3).astype(np.float32)
        x = rng.standard_normal((rows, xcols)).astype(np.float32)
        dy = rng.standard_normal((rows, ycols)).astype(np.float32)
        y, dx, g = R.layer(kind, n_in, n_out, p, x, ycols, dy)
        out[f"{name}_p"], out[f"{name}_x"], out[f"{name}_dy"] = p, x, dy
        out[f"{name}_y"], out[f"{name}_dx"], out[f"{name}_g"] = y, dx, g
    x = rng.standard_normal((23, 8)).astype(np.float32)
    dy = rng.standard_normal((23, 8)).astype(np.float32)
    for name, kind in [("relu", R.RELU), ("softmax", R.SOFTMAX), ("softmax_ce", R.SOFTMAX_CE)]:
        y, dx, _ = R.layer(kind, 0, 0, None, x, 8, dy)
        out[f"{name}_x"], out[f"{name}_dy"], out[f"{name}_y"], out[f"{name}_dx"] = x, dy, y, dx

    # whole models (nn.h:467-542)
    for name, net, cols, ycols in [
        ("mlp_c2_policy", R.fc_net([32, 64, 64, 8], R.SOFTMAX), 32, 8),
        ("mlp_value", R.fc_net([32, 64, 32, 1]), 32, 1),
        ("mlp_conv_policy", R.conv_net([4, 16, 8, 1], R.SOFTMAX), 32, 8),
        ("mlp_pg_policy", R.fc_net([32, 24, 12, 8], R.SOFTMAX_CE), 32, 8),
    ]:
        p = R.init_params(net, 3)
        p = (p + rng.standard_normal(p.size).astype(np.float32) * 0.05).astype(np.float32)
        x = (rng.integers(0, 9, (21, cols)) / 8.0).astype(np.float32)
        dy = rng.standard_normal((21, ycols)).astype(np.float32)
        g, o = R.model_gradient(net, p, x, dy)
        out[f"{name}_layers"] = np.array(net.layers, np.int32)
        out[f"{name}_p"], out[f"{name}_x"], out[f"{name}_dy"] = p, x, dy
        out[f"{name}_g"], out[f"{name}_out"] = g, o
    init = R.init_params(R.fc_net([32, 64, 64, 8], R.SOFTMAX), 5)
    out["init_dense_std"] = np.float32(init[:32 * 64].std())
    initc = R.init_params(R.conv_net([4, 128, 64, 1]), 5)
    out["init_conv_std_l1"] = np.float32(initc[:4 * 128].std())

    # loss-gradient rules (rl.h:45-74; policy_gradient.h:47-85)
    P = rng.random((40, 8)).astype(np.float32) + 0.05
    P /= P.sum(1, keepdims=True)
    PO = rng.random((40, 8)).astype(np.float32) + 0.05
    PO /= PO.sum(1, keepdims=True)
    ch = rng.integers(0, 8, 40).astype(np.int32)
    adv = rng.standard_normal(40).astype(np.float32) * 2
    out["loss_p"], out["loss_pold"], out["loss_choice"], out["loss_adv"] = P, PO, ch, adv
    out["loss_softmax_log"] = np.stack([R.action_gradient(0, P[i], PO[i], ch[i], adv[i]) for i in range(40)])
    out["loss_clipped"] = np.stack([R.action_gradient(1, P[i], PO[i], ch[i], adv[i]) for i in range(40)])
    kl, beta = R.kl_loss(P, PO, ch, adv, 1e-9, 1.0)
    out["loss_kl_beta1"], out["loss_kl_beta_next"] = kl, np.float32(beta)
    kl2, beta2 = R.kl_loss(P, P, ch, adv, 1e-9, 0.05)
    out["loss_kl_same"], out["loss_kl_same_beta_next"] = kl2, np.float32(beta2)

    # optimizers (nn.h:616-698)
    p0 = rng.standard_normal(50).astype(np.float32)
    gs = rng.standard_normal((5, 50)).astype(np.float32)
    out["opt_p0"], out["opt_grads"] = p0, gs
    out["opt_sgd"] = R.opt_steps(R.SGD, 1e-2, 0.0, p0, gs)
    out["opt_sgd_wd"] = R.opt_steps(R.SGD, 1e-2, 1e-3, p0, gs)
    out["opt_momentum"] = R.opt_steps(R.MOMENTUM, 1e-2, 0.0, p0, gs)
    out["opt_adam"] = R.opt_steps(R.ADAM, 1e-2, 0.0, p0, gs)

    # environment dynamics with forced actions (bin_packing.h:46-107; rl.h:325-349)
    acts = rng.integers(0, 8, 400).astype(np.int32)
    out["env_forced_actions"] = acts
    out["env_forced_steps"] = R.env_forced(99, acts)
    acts2 = (np.arange(300) % 8).astype(np.int32)
    out["env_rr_actions"] = acts2
    out["env_rr_steps"] = R.env_forced(3, acts2)

    # deep_agent known answer (deep_agent.cc; weights.20) on a short run
    w20 = np.fromfile("/root/reference/apps/bin_packing/weights.20", dtype=np.float32)
    mean, nsteps = R.eval_argmax(2021, R.conv_net([4, 128, 64, 1]), w20, 300)
    # the trained weights themselves (raw flat fp32 checkpoint format, deep_agent.cc:21-23): a
    # 35 844-byte DATA fixture of the reference, needed for the known-answer eval on the GPU box
    out["weights20"] = w20
    out["deep_agent_mean_300ep_seed2021"] = np.float64(mean)
    out["deep_agent_steps_300ep_seed2021"] = np.int64(nsteps)
    return out


def main():
    """No arguments: regenerate everything. With arguments: only the named trainer cases."""
    os.makedirs(refcases.GOLDEN_DIR, exist_ok=True)
    only = set(sys.argv[1:])
    if not only:
        np.savez_compressed(os.path.join(refcases.GOLDEN_DIR, "units.npz"), **units())
        print("units.npz written")
    for c in refcases._cases():
        if only and c["name"] not in only:
            continue
        arrs = refcases.generate_case(c)
        path = os.path.join(refcases.GOLDEN_DIR, f"train_{c['name']}.npz")
        np.savez_compressed(path, **arrs)
        print(c["name"], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
