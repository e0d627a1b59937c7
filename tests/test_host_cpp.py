"""The C++20 host mirror of the reference classes (dependence_free_rl_b200/host): builds with g++
against include/dfrl.h only, and on a GPU produces exactly what the Python binding of the same C
ABI produces for the same configuration."""
import os
import subprocess

import numpy as np
import pytest

import refcases

ROOT = refcases.ROOT
HOST = os.path.join(ROOT, "dependence_free_rl_b200", "host")
BINS = ["ppo_training", "ppo2_training", "ac_training", "pg_training", "deep_agent", "heuristic_agent", "host_api_test", "flagstore_test"]


def _build():
    subprocess.check_call(["make", "-C", HOST, "-j8"], stdout=subprocess.DEVNULL)


def test_host_mirror_builds_and_keeps_reference_signatures():
    _build()
    for b in BINS:
        assert os.path.exists(os.path.join(HOST, ".out", b)), b
    rl = open(os.path.join(HOST, "xylo", "rl.h")).read()
    pg = open(os.path.join(HOST, "xylo", "policy_gradient.h")).read()
    nn = open(os.path.join(HOST, "xylo", "nn.h")).read()
    # signatures the north star names (reference rl.h:163-170, 317-392; policy_gradient.h:92-317; nn.h:20-33)
    for sig in ["virtual void apply(const A &action, std::size_t id) = 0;", "virtual S view(std::size_t id) const = 0;",
                "virtual void reset(std::size_t id) = 0;", "virtual A react(const S &state) const = 0;",
                "explicit agent(const policy<A, S> &p, environment<A, S> &env, replay_buffer<A, S> &rb, std::size_t id = 0)",
                "void play_steps(std::size_t n)", "void play_one_episode()", "virtual void learn() = 0;"]:
        assert sig in rl, sig
    for cls in ["policy_gradient_learner", "actor_critic_learner", "ppo_learner", "kl_ppo_learner",
                "policy_gradient_policy", "policy_gradient_deterministic_policy"]:
        assert f"class {cls}" in pg, cls
    # the reference's specialisation hook (policy_gradient.h:187) and the rule check of agent::game_over / get_reward
    assert "virtual void optimize_action(matrix_view" in pg and "const std::vector<A> &" in pg
    assert "void verify_rules()" in rl
    # trajectory types and replay_buffer::sample_td (reference rl.h:111-234): a read-back of the device records
    for sig in ["template <typename A, typename S> struct transition {", "template <typename A, typename S> struct trajectory {",
                "template <typename A, typename S> class td {", "std::vector<td<A, S>> sample_td()",
                "float total_rewards(const std::vector<td<A, S>> &experience)"]:
        assert sig in rl, sig
    for sig in ["virtual matrix forward(matrix_view t) = 0;", "virtual matrix backward(matrix_view input, matrix_view loss) = 0;",
                "virtual vector gradient(matrix_view input, matrix_view backprop) = 0;",
                "void step(matrix_view input, const loss_grad_func &loss_grad)",
                "sgd_optimizer(model &m, float rate, float weight_decay = 0.0f)",
                "adam_optimizer(model &m, float rate, float beta1 = 0.9, float beta2 = 0.999)"]:
        assert sig in nn, sig
    # the host side reaches CUDA only through the C ABI
    for dirpath, _, files in os.walk(HOST):
        for f in files:
            if f.endswith((".h", ".cc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "cuda_runtime" not in txt and "<<<" not in txt, f


def test_flagstore_and_runtime_problem_definition():
    """xeno::flagstore (reference xeno/configuration.h:17-118) and bp::configure: the reference's constexpr
    num_bins / capacity / item shapes as launch parameters. No device needed."""
    _build()
    out = subprocess.run([os.path.join(HOST, ".out", "flagstore_test")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


def test_own_xmake_builds_the_host_targets(tmp_path):
    """The repo's own build tool (tools/xmake/xmake.cc, C++ like the reference's build/xmake.cc; target
    schema of build/xmake.cc:92-103 + the `cuda` rule): stage-0 bootstrap, a dry-run plan that compiles the
    .cu sources with nvcc for sm_100a and links the host main against the device library only, and an
    mtime-based no-op on the second run."""
    xm = subprocess.run([os.path.join(ROOT, "tools", "xmake", "bootstrap.sh")], capture_output=True, text=True, timeout=300)
    assert xm.returncode == 0, xm.stdout + xm.stderr
    exe = xm.stdout.strip().splitlines()[-1]
    pkg = os.path.join(HOST, "apps", "bin_packing")
    plan = subprocess.run([exe, "-n", "-B", "ppo_training"], cwd=pkg, capture_output=True, text=True, timeout=60)
    assert plan.returncode == 0, plan.stdout + plan.stderr
    lines = plan.stdout.strip().splitlines()
    cu = [l for l in lines if l.startswith("nvcc") and " -c " in l]
    assert len(cu) == 8 and all("arch=compute_100a,code=sm_100a" in l and "-lineinfo" in l for l in cu), plan.stdout
    assert any("-shared" in l and "libdfrl_b200.so" in l for l in lines)
    assert "-std=c++20" in lines[-2] and "ppo_training.cc" in lines[-2]
    assert "-ldfrl_b200" in lines[-1] and "cudart" not in lines[-1] and lines[-1].rstrip().endswith("ppo_training")
    # a real build of a host-only test target (the device library is linked, not rebuilt, when it is current)
    tests_pkg = os.path.join(HOST, "tests")
    lib = os.path.join(ROOT, "dependence_free_rl_b200", "csrc", ".out", "libdfrl_b200.so")
    if os.path.exists(lib):
        out = subprocess.run([exe, "flagstore_test"], cwd=tests_pkg, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout + out.stderr
        again = subprocess.run([exe, "flagstore_test"], cwd=tests_pkg, capture_output=True, text=True, timeout=60)
        assert again.returncode == 0 and again.stdout.strip() == "", again.stdout
        run = subprocess.run([os.path.join(tests_pkg, ".out", "flagstore_test")], capture_output=True, text=True, timeout=60)
        assert run.returncode == 0 and run.stdout.strip().endswith("OK")
    bad = subprocess.run([exe, "no_such_target"], cwd=pkg, capture_output=True, text=True, timeout=60)
    assert bad.returncode != 0 and "no target" in bad.stderr


@pytest.mark.gpu
def test_host_api_matches_python_binding(D, ctx):
    _build()
    out = subprocess.run([os.path.join(HOST, ".out", "host_api_test")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    lines = {l.split()[0]: l.split()[1:] for l in out.stdout.splitlines() if l and not l.startswith("OK")}
    assert "FAIL" not in out.stdout
    got = dict(zip(lines["ppo"][0::2], lines["ppo"][1::2]))
    # a user's optimize_action override was called once per learn() with every start row; an agent
    # with a different game_over / get_reward rule was rejected
    assert lines["hook"][:4] == ["calls", "2", "rows", "384"], lines["hook"]
    assert lines["odd_agent"] == ["rejected", "1"]
    # the same run through the Python binding
    n, T, iters = 4096, 4, 5
    rs = np.float32(32.0) / np.float32(n * T)
    policy = D.Model(ctx, D.fc_layers([32, 64, 64, 8], D.SOFTMAX), 32)
    value = D.Model(ctx, D.fc_layers([32, 64, 64, 1]), 32)
    policy.init_parameters(1234)
    value.init_parameters(1235)
    env = D.Environment(ctx, n, seed=1234)
    tr = D.Trainer(ctx, env, policy, value, algo=D.PPO, work=T, policy_lr=float(np.float32(1e-4) * rs),
                   value_lr=float(np.float32(1e-5) * rs))
    tr.iterate(iters)
    s = tr.stats()
    def checksum(p):  # the C++ side's left-to-right double sum of p[i] * (i % 7 + 1)
        acc = 0.0
        for i, x in enumerate(p.astype(np.float64)):
            acc += x * float(i % 7 + 1)
        return acc
    assert int(got["env_steps"]) == s["env_steps"] == n * T * iters
    assert int(got["episodes"]) == s["episodes"]
    assert float(got["reward_sum"]) == s["reward_sum"]
    assert float(got["policy_sum"]) == checksum(policy.parameters())   # bit-identical parameters
    assert float(got["value_sum"]) == checksum(value.parameters())
    tr.close(); env.close(); policy.close(); value.close()


@pytest.mark.gpu
def test_deep_agent_cpp_known_answer(tmp_path):
    # deep_agent.cc with weights.20: 26.553 +- 0.028 per 10 000-episode round (reference deep.log)
    _build()
    U = np.load(os.path.join(refcases.GOLDEN_DIR, "units.npz"))
    wf = tmp_path / "weights.20"
    U["weights20"].astype("<f4").tofile(wf)
    out = subprocess.run([os.path.join(HOST, ".out", "deep_agent"), str(wf), "8192", "4"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    mean = float(out.stdout.split()[1])
    assert 26.45 < mean < 26.65, out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("rule,lo,hi", [("minwaste", 26.45, 26.65), ("firstfit", 25.7, 25.95), ("bestfit", 25.7, 25.95),
                                        ("random", 11.3, 11.9)])
def test_heuristic_agents_reach_the_reference_levels(rule, lo, hi):
    """firstfit_agent.cc / bestfit_agent.cc / minwaste_agent.cc / random_agent.cc through the host mirror: rule-based
    policies behind xylo::policy (device_rule), whole episodes on the device; mean reward per episode at the
    reference's logged levels (minwaste.log: 26.553 +- 0.009)."""
    _build()
    out = subprocess.run([os.path.join(HOST, ".out", "heuristic_agent"), rule, "16384", "4", "2"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    rounds = [float(l.split()[3]) for l in out.stdout.splitlines() if l.startswith("round")]
    assert len(rounds) == 2 and all(lo < r < hi for r in rounds), out.stdout
    react = [l for l in out.stdout.splitlines() if l.startswith("react(")]
    assert react and (rule == "random" or react[0].endswith("bin 0")), out.stdout   # an empty board: the first bin


@pytest.mark.gpu
@pytest.mark.parametrize("binary,args", [("ppo_training", ["4096", "6", "5"]), ("ppo_training", ["4096", "6", "5", "c2"]),
                                         ("ppo_training", ["--num_bins=16", "1024", "3", "2", "c2"]),   # run-time bins: layered path
                                         ("ppo_training", ["-b", "5", "--capacity", "16", "512", "3", "0"]),
                                         ("ppo2_training", ["1024", "5", "4"]),   # KL-PPO: beta on the device
                                         ("ac_training", ["2048", "4"]),
                                         ("pg_training", ["256", "3"])])
def test_trainer_mains_run(binary, args):
    _build()
    out = subprocess.run([os.path.join(HOST, ".out", binary)] + args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    last = out.stdout.strip().splitlines()[-1].split()
    kv = dict(zip(last[0::2], last[1::2]))
    assert int(kv["env_steps"]) > 0 and float(kv["reward_sum"]) + int(kv["episodes"]) == int(kv["env_steps"])


@pytest.mark.gpu
@pytest.mark.parametrize("nets", ["dense", "conv"])
def test_p2p_gradient_exchange_two_ranks(nets):
    """Fused peer-memory exchange (dfrl_p2p_*) on 2 GPUs: bit-identical parameters on both ranks and
    the NCCL path's result -- for the dense nets (tcgen05 learner kernels) and for the reference's conv1d policy
    on the table path (conv_table.cuh: the same gradient tail). Needs 2 visible GPUs (`gpurun --gpus 2`); skipped
    on a 1-GPU box."""
    import json
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29571" if nets == "dense" else "29572", os.path.join(ROOT, "tests", "p2p_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, P2P_NETS=nets))
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["ok"] and d["identical_across_ranks"], d
