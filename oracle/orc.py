"""ctypes binding of oracle/liboracle.so (dfrl_oracle.c) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this module.  The product package (dependence_free_rl_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

DENSE, CONV1D, RELU, SOFTMAX, SOFTMAX_CE = 0, 1, 2, 3, 4
REINFORCE, ACTOR_CRITIC, PPO, KL_PPO = 0, 1, 2, 3
SGD, MOMENTUM, ADAM = 0, 1, 2
LOSS_SOFTMAX_LOG, LOSS_CLIPPED, LOSS_KL = 0, 1, 2
HEUR_RANDOM, HEUR_FIRSTFIT, HEUR_BESTFIT, HEUR_MINWASTE = 0, 1, 2, 3
MAX_LAYERS = 16


class NetC(C.Structure):
    _fields_ = [("n", C.c_int), ("kind", C.c_int * MAX_LAYERS), ("in_", C.c_int * MAX_LAYERS),
                ("out", C.c_int * MAX_LAYERS), ("input_cols", C.c_int),
                # shared-trunk extension: explicit parameter offsets into one flat vector (0 = sequential)
                ("n_params", C.c_int), ("poff", C.c_int * MAX_LAYERS)]


class EnvCfgC(C.Structure):
    _fields_ = [("n_bins", C.c_int), ("cap_w", C.c_int), ("cap_h", C.c_int),
                ("item_w", C.c_int * 2), ("item_h", C.c_int * 2), ("p_shape1", C.c_double)]


class TrainCfgC(C.Structure):
    _fields_ = [("algo", C.c_int), ("work", C.c_int), ("gamma", C.c_float), ("lambda_", C.c_float),
                ("epochs", C.c_int), ("kl_target", C.c_float), ("policy_opt", C.c_int),
                ("value_opt", C.c_int), ("policy_lr", C.c_float), ("value_lr", C.c_float),
                ("policy_wd", C.c_float), ("value_wd", C.c_float), ("adam_beta1", C.c_float),
                ("adam_beta2", C.c_float)]


def build(force=False):
    """Compiles the oracle (gcc, seconds). Safe to call repeatedly."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "dfrl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


_libs = {}


def lib(f64=False):
    """f64: False = float accumulators (the reference's), True = double accumulators,
    "mt" = double accumulators + OpenMP over rows / outputs (same bits as True, any thread count)."""
    key = "64mt" if f64 == "mt" else "64" if f64 else ""
    if key not in _libs:
        build()
        l = C.CDLL(os.path.join(_HERE, f"liboracle{key}.so"))
        l.orc_canonical.restype = C.c_double
        l.orc_minstd_next.restype = C.c_uint32
        l.orc_minstd_seed.restype = C.c_uint32
        l.orc_kl_next_beta.restype = C.c_float
        _libs[key] = l
    return _libs[key]


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class Net:
    """[(kind, in, out), ...] + input width, same vocabulary as oracle.ref.Net."""

    def __init__(self, layers, input_cols):
        self.layers = list(layers)
        self.input_cols = input_cols
        self.c = NetC()
        self.c.n = len(layers)
        for i, (k, a, b) in enumerate(layers):
            self.c.kind[i], self.c.in_[i], self.c.out[i] = k, a, b
        self.c.input_cols = input_cols

    def param_count(self):
        return lib().orc_net_param_count(C.byref(self.c))

    def with_offsets(self, offsets, n_params):
        """Explicit parameter offset per layer (None for activation layers) into ONE flat vector of
        n_params floats: two nets that give their first layers the same offsets share a trunk."""
        net = Net(self.layers, self.input_cols)
        net.c.n_params = n_params
        for i, o in enumerate(offsets):
            net.c.poff[i] = 0 if o is None else o
        return net

    def output_cols(self):
        return lib().orc_net_output_cols(C.byref(self.c))


def fc_net(dims, last=None):
    layers = []
    for i in range(len(dims) - 1):
        layers.append((DENSE, dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append((RELU, 0, 0))
    if last is not None:
        layers.append((last, 0, 0))
    return Net(layers, dims[0])


def conv_net(chans, points, last=None):
    layers = []
    for i in range(len(chans) - 1):
        layers.append((CONV1D, chans[i], chans[i + 1]))
        if i < len(chans) - 2:
            layers.append((RELU, 0, 0))
    if last is not None:
        layers.append((last, 0, 0))
    return Net(layers, chans[0] * points)


def env_cfg(n_bins=8):
    c = EnvCfgC()
    lib().orc_env_cfg_default(C.byref(c))
    c.n_bins = n_bins
    return c


# ---- RNG ----
class Minstd:
    def __init__(self, seed):
        self.s = C.c_uint32(lib().orc_minstd_seed(C.c_uint32(seed)))

    def next(self):
        return lib().orc_minstd_next(C.byref(self.s))

    def canonical(self):
        return lib().orc_canonical(C.byref(self.s))

    def bernoulli(self, p):
        return lib().orc_bernoulli(C.byref(self.s), C.c_double(p))


def discrete(w, u):
    w = f32(w)
    return lib().orc_discrete(_p(w), w.size, C.c_double(u))


def argmax(w):
    w = f32(w)
    return lib().orc_argmax(_p(w), w.size)


# ---- env ----
def env_reset_all(cfg, n, first_item):
    st = np.zeros((2 * cfg.n_bins + 2, n), dtype=np.int8)
    fi = u8(first_item)
    lib().orc_env_reset_all(C.byref(cfg), _p(st), n, _p(fi))
    return st


def env_step(cfg, state, actions, next_item, want_terminal=False):
    n = state.shape[1]
    done = np.zeros(n, dtype=np.uint8)
    term = np.zeros_like(state) if want_terminal else None
    a, it = u8(actions), u8(next_item)
    lib().orc_env_step(C.byref(cfg), _p(state), n, _p(a), _p(it), _p(done), _p(term))
    return done, term


def obs_encode(state, n_bins, cap_w=8, cap_h=8):
    state = np.ascontiguousarray(state, dtype=np.int8)
    rows = state.shape[1]
    obs = np.zeros((rows, 4 * n_bins), dtype=np.float32)
    lib().orc_obs_encode(_p(state), rows, rows, n_bins, cap_w, cap_h, _p(obs))
    return obs


def heuristic_react(cfg, state, env, kind, u=0.0):
    return lib().orc_heuristic_react(C.byref(cfg), _p(state), state.shape[1], env, kind,
                                     C.c_double(u))


# ---- layers / model ----
def dense_forward(params, n_in, n_out, x, f64=False):
    x, params = f32(x), f32(params)
    y = np.zeros((x.shape[0], n_out), dtype=np.float32)
    lib(f64).orc_dense_forward(_p(params), n_in, n_out, _p(x), x.shape[0], _p(y))
    return y


def dense_backward(params, n_in, n_out, dy, f64=False):
    dy, params = f32(dy), f32(params)
    dx = np.zeros((dy.shape[0], n_in), dtype=np.float32)
    lib(f64).orc_dense_backward(_p(params), n_in, n_out, _p(dy), dy.shape[0], _p(dx))
    return dx


def dense_gradient(n_in, n_out, x, dy, f64=False):
    x, dy = f32(x), f32(dy)
    g = np.zeros((n_in + 1) * n_out, dtype=np.float32)
    lib(f64).orc_dense_gradient(n_in, n_out, _p(x), _p(dy), x.shape[0], _p(g))
    return g


def relu_forward(x):
    x = f32(x)
    y = np.zeros_like(x)
    lib().orc_relu_forward(_p(x), C.c_size_t(x.size), _p(y))
    return y


def relu_backward(x, dy):
    x, dy = f32(x), f32(dy)
    dx = np.zeros_like(x)
    lib().orc_relu_backward(_p(x), _p(dy), C.c_size_t(x.size), _p(dx))
    return dx


def softmax_forward(x):
    x = f32(x)
    y = np.zeros_like(x)
    lib().orc_softmax_forward(_p(x), x.shape[0], x.shape[1], _p(y))
    return y


def softmax_backward(x, dy):
    x, dy = f32(x), f32(dy)
    dx = np.zeros_like(x)
    lib().orc_softmax_backward(_p(x), _p(dy), x.shape[0], x.shape[1], _p(dx))
    return dx


def net_eval(net, params, x, f64=False):
    x, params = f32(x), f32(params)
    y = np.zeros((x.shape[0], net.output_cols()), dtype=np.float32)
    lib(f64).orc_net_eval(C.byref(net.c), _p(params), _p(x), x.shape[0], _p(y))
    return y


def net_forward_gradient(net, params, x, dy, f64=False):
    x, params, dy = f32(x), f32(params), f32(dy)
    g = np.zeros(params.size, dtype=np.float32)
    out = np.zeros((x.shape[0], net.output_cols()), dtype=np.float32)
    lib(f64).orc_net_forward_gradient(C.byref(net.c), _p(params), _p(x), x.shape[0], _p(dy), _p(g),
                                      _p(out))
    return g, out


# ---- losses ----
def loss_grad(kind, probs, actions, adv, p_old=None, beta=0.0):
    probs, adv = f32(probs), f32(adv)
    a = u8(actions)
    po = None if p_old is None else f32(p_old)
    out = np.zeros_like(probs)
    lib().orc_loss_grad(kind, _p(probs), _p(a), _p(adv), _p(po), C.c_float(beta), probs.shape[0],
                        probs.shape[1], _p(out))
    return out


def kl_next_beta(probs, p_old, d_targ, beta):
    probs, p_old = f32(probs), f32(p_old)
    return lib().orc_kl_next_beta(_p(probs), _p(p_old), probs.shape[0], probs.shape[1],
                                  C.c_float(d_targ), C.c_float(beta))


# ---- returns / gae ----
def returns(done, length, gamma):
    done = u8(done)
    L, n = done.shape
    ln = None if length is None else np.ascontiguousarray(length, dtype=np.int32)
    g = np.zeros((L, n), dtype=np.float32)
    acc = np.zeros(2, dtype=np.float64)
    lib().orc_returns(_p(done), _p(ln), n, L, C.c_float(gamma), _p(g), _p(acc))
    return g, acc


def gae(done, v_start, v_end, gamma, lam):
    done, v_start, v_end = u8(done), f32(v_start), f32(v_end)
    T, n = done.shape
    tg = np.zeros((T, n), dtype=np.float32)
    adv = np.zeros((T, n), dtype=np.float32)
    lib().orc_gae(_p(done), _p(v_start), _p(v_end), n, T, C.c_float(gamma), C.c_float(lam), _p(tg),
                  _p(adv))
    return tg, adv


# ---- optimizers ----
def opt_step(kind, params, grad, state, lr, wd=0.0, b1=0.9, b2=0.999, t=1.0):
    """In place on params/state (float32 arrays)."""
    assert params.dtype == np.float32 and params.flags.c_contiguous
    g = f32(grad)
    lib().orc_opt_step(kind, _p(params), _p(g), _p(state), params.size, C.c_float(lr),
                       C.c_float(wd), C.c_float(b1), C.c_float(b2), C.c_float(t))


def train_cfg(algo, work, gamma=0.99, lam=0.95, epochs=4, kl_target=1e-9, policy_opt=SGD,
              value_opt=SGD, policy_lr=1e-4, value_lr=1e-5, policy_wd=0.0, value_wd=0.0,
              b1=0.9, b2=0.999):
    c = TrainCfgC()
    c.algo, c.work, c.gamma, c.lambda_, c.epochs, c.kl_target = algo, work, gamma, lam, epochs, kl_target
    c.policy_opt, c.value_opt = policy_opt, value_opt
    c.policy_lr, c.value_lr, c.policy_wd, c.value_wd = policy_lr, value_lr, policy_wd, value_wd
    c.adam_beta1, c.adam_beta2 = b1, b2
    return c


def opt_state_size(kind, n):
    return {SGD: 0, MOMENTUM: n, ADAM: 2 * n}[kind]


class Learner:
    """Holds params/optimizer state across orc_learn calls (the reference's learner objects)."""

    def __init__(self, tcfg, ecfg, pnet, pparams, vnet=None, vparams=None, kl_beta0=1.0, f64=False):
        self.t, self.e, self.pnet, self.vnet = tcfg, ecfg, pnet, vnet
        self.pparams = f32(pparams).copy()
        # shared-trunk nets (Net.with_offsets): pass vparams=None with a value net -> both nets
        # address the SAME flat vector, so the critic step moves the trunk the actor step then sees
        shared = vnet is not None and vparams is None
        self.vparams = self.pparams if shared else (None if vparams is None else f32(vparams).copy())
        self.pstate = np.zeros(max(1, opt_state_size(tcfg.policy_opt, self.pparams.size)), np.float32)
        self.vstate = np.zeros(max(1, opt_state_size(tcfg.value_opt, 0 if self.vparams is None else self.vparams.size)), np.float32)
        self.p_t, self.v_t = C.c_float(1.0), C.c_float(1.0)
        self.kl_beta = C.c_float(kl_beta0)
        self.f64 = f64

    def learn(self, rec_state, final_state, action, done, p_old, length=None):
        rec_state = np.ascontiguousarray(rec_state, dtype=np.int8)
        final_state = np.ascontiguousarray(final_state, dtype=np.int8)
        action, done, p_old = u8(action), u8(done), f32(p_old)
        L, n = action.shape
        ln = None if length is None else np.ascontiguousarray(length, dtype=np.int32)
        adv = np.zeros((L, n), np.float32)
        tg = np.zeros((L, n), np.float32)
        epochs = 1 if self.t.algo in (REINFORCE, ACTOR_CRITIC) else self.t.epochs
        pg = np.zeros((epochs, self.pparams.size), np.float32)
        vg = np.zeros(1 if self.vparams is None else self.vparams.size, np.float32)
        rc = lib(self.f64).orc_learn(
            C.byref(self.t), C.byref(self.e), n, L, _p(rec_state), _p(final_state), _p(action),
            _p(done), _p(ln), _p(p_old), C.byref(self.pnet.c), _p(self.pparams), _p(self.pstate),
            C.byref(self.p_t), None if self.vnet is None else C.byref(self.vnet.c),
            _p(self.vparams), _p(self.vstate), C.byref(self.v_t), C.byref(self.kl_beta), _p(adv),
            _p(tg), _p(vg), _p(pg))
        if rc:
            raise RuntimeError("orc_learn failed")
        return {"adv": adv, "targets": tg, "policy_grads": pg, "value_grad": vg}


def rollout(ecfg, state, pnet, pparams, L, mode, items, forced=None, u=None):
    """mode 0 sample(u) / 1 argmax / 2 forced. state is updated in place."""
    n = state.shape[1]
    B = ecfg.n_bins
    items = u8(items)
    fo = None if forced is None else u8(forced)
    uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    rs = np.zeros((L, 2 * B + 2, n), np.int8)
    ra = np.zeros((L, n), np.uint8)
    rd = np.zeros((L, n), np.uint8)
    rp = np.zeros((L, n, B), np.float32)
    pp = f32(pparams)
    lib().orc_rollout(C.byref(ecfg), _p(state), n, L, C.byref(pnet.c), _p(pp), mode, _p(items),
                      _p(fo), _p(uu), _p(rs), _p(ra), _p(rd), _p(rp))
    return {"state": rs, "action": ra, "done": rd, "probs": rp}
