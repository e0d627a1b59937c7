"""ctypes binding of oracle/_ref/libdfrl_ref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library is the UNMODIFIED reference (beehover/dependence_free_rl) compiled by
oracle/Makefile plus oracle/ref_harness.cc.  Only tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdfrl_ref.so")

DENSE, CONV1D, RELU, SOFTMAX, SOFTMAX_CE = 0, 1, 2, 3, 4
REINFORCE, ACTOR_CRITIC, PPO, KL_PPO = 0, 1, 2, 3
SGD, MOMENTUM, ADAM = 0, 1, 2

NB = 8  # bp::num_bins of the compiled reference (bin_packing.h:12)

STEP_DTYPE = np.dtype([
    ("iter", "<i4"), ("env", "<i4"), ("t", "<i4"), ("action", "<i4"), ("done", "<i4"),
    ("reward", "<f4"), ("sbins", "<i4", (NB, 2)), ("sitem", "<i4", (2,)),
    ("ebins", "<i4", (NB, 2)), ("eitem", "<i4", (2,)), ("item_after", "<i4", (2,)),
    ("p_old", "<f4", (NB,)),
])
ROW_DTYPE = np.dtype([("iter", "<i4"), ("env", "<i4"), ("t", "<i4"), ("frozen", "<i4"),
                      ("advantage", "<f4")])

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` in the dev container")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_train.restype = C.c_double
        _lib.ref_eval_argmax.restype = C.c_double
        _lib.ref_last_error.restype = C.c_char_p
        assert _lib.ref_step_rec_size() == STEP_DTYPE.itemsize
        assert _lib.ref_row_rec_size() == ROW_DTYPE.itemsize
        assert _lib.ref_num_bins() == NB
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class Net:
    """A layer list in the reference's vocabulary: [(kind, in, out), ...]."""

    def __init__(self, layers):
        self.layers = list(layers)
        self.kinds = np.array([l[0] for l in layers], dtype=np.int32)
        self.ins = np.array([l[1] for l in layers], dtype=np.int32)
        self.outs = np.array([l[2] for l in layers], dtype=np.int32)

    @property
    def n(self):
        return len(self.layers)

    def args(self):
        return self.n, _ip(self.kinds), _ip(self.ins), _ip(self.outs)

    def param_count(self):
        return lib().ref_param_count(*self.args())


def fc_net(dims, last=None):
    """Dense MLP dims[0]->...->dims[-1] with relu between and optional final layer kind."""
    layers = []
    for i in range(len(dims) - 1):
        layers.append((DENSE, dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append((RELU, 0, 0))
    if last is not None:
        layers.append((last, 0, 0))
    return Net(layers)


def conv_net(chans, last=None):
    layers = []
    for i in range(len(chans) - 1):
        layers.append((CONV1D, chans[i], chans[i + 1]))
        if i < len(chans) - 2:
            layers.append((RELU, 0, 0))
    if last is not None:
        layers.append((last, 0, 0))
    return Net(layers)


def engine_draw(seed, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().ref_engine_draw(C.c_uint(seed), n, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def init_params(net, seed):
    p = np.zeros(net.param_count(), dtype=np.float32)
    lib().ref_init_params(C.c_uint(seed), *net.args(), _fp(p))
    return p


def model_eval(net, params, x, ycols):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros((x.shape[0], ycols), dtype=np.float32)
    rc = lib().ref_model_eval(*net.args(), _fp(np.ascontiguousarray(params, dtype=np.float32)),
                              _fp(x), x.shape[0], x.shape[1], _fp(y), ycols)
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return y


def model_gradient(net, params, x, dy):
    x = np.ascontiguousarray(x, dtype=np.float32)
    dy = np.ascontiguousarray(dy, dtype=np.float32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    grad = np.zeros(params.size, dtype=np.float32)
    out = np.zeros_like(dy)
    rc = lib().ref_model_gradient(*net.args(), _fp(params), _fp(x), x.shape[0], x.shape[1],
                                  _fp(dy), dy.shape[1], _fp(grad), _fp(out))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return grad, out


def layer(kind, n_in, n_out, params, x, ycols, dy=None):
    """Returns (y, dx, grad) of one reference layer."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros((x.shape[0], ycols), dtype=np.float32)
    dx = np.zeros_like(x) if dy is not None else None
    npar = (n_in + 1) * n_out if kind in (DENSE, CONV1D) else 0
    grad = np.zeros(npar, dtype=np.float32) if (dy is not None and npar) else None
    p = None if params is None else np.ascontiguousarray(params, dtype=np.float32)
    d = None if dy is None else np.ascontiguousarray(dy, dtype=np.float32)
    rc = lib().ref_layer(kind, n_in, n_out, _fp(p), _fp(x), x.shape[0], x.shape[1], _fp(d), ycols,
                         _fp(y), _fp(dx), _fp(grad))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return y, dx, grad


def action_gradient(kind, p, p_old, choice, adv):
    p = np.ascontiguousarray(p, dtype=np.float32)
    po = np.ascontiguousarray(p_old, dtype=np.float32)
    out = np.zeros(NB, dtype=np.float32)
    rc = lib().ref_action_gradient(kind, _fp(p), _fp(po), int(choice), C.c_float(adv), _fp(out))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return out


def kl_loss(p, p_old, choices, adv, d_targ, beta):
    p = np.ascontiguousarray(p, dtype=np.float32)
    po = np.ascontiguousarray(p_old, dtype=np.float32)
    ch = np.ascontiguousarray(choices, dtype=np.int32)
    a = np.ascontiguousarray(adv, dtype=np.float32)
    out = np.zeros_like(p)
    b = C.c_float(beta)
    rc = lib().ref_kl_loss(p.shape[0], _fp(p), _fp(po), _ip(ch), _fp(a), C.c_float(d_targ),
                           C.byref(b), _fp(out))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return out, b.value


def discrete_sample(seed, w, count):
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.zeros(count, dtype=np.int32)
    lib().ref_discrete_sample(C.c_uint(seed), _fp(w), w.size, count, _ip(out))
    return out


def argmax(w):
    w = np.ascontiguousarray(w, dtype=np.float32)
    return lib().ref_argmax(_fp(w), w.size)


def opt_steps(kind, lr, wd, params0, grads):
    params0 = np.ascontiguousarray(params0, dtype=np.float32)
    grads = np.ascontiguousarray(grads, dtype=np.float32)
    k, n = grads.shape
    out = np.zeros((k, n), dtype=np.float32)
    rc = lib().ref_opt_steps(kind, C.c_float(lr), C.c_float(wd), n, _fp(params0), k, _fp(grads),
                             _fp(out))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return out


def _steps():
    n = lib().ref_steps_count()
    a = np.zeros(n, dtype=STEP_DTYPE)
    if n:
        lib().ref_steps_copy(a.ctypes.data_as(C.c_void_p))
    return a


def _rows():
    n = lib().ref_rows_count()
    a = np.zeros(n, dtype=ROW_DTYPE)
    if n:
        lib().ref_rows_copy(a.ctypes.data_as(C.c_void_p))
    return a


def env_forced(seed, actions):
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    n = lib().ref_env_forced(C.c_uint(seed), actions.size, _ip(actions))
    if n < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return _steps()


def train(algo, seed, n_envs, work, iters, policy, pparams, plr, value=None, vparams=None,
          vlr=0.0, popt=SGD, vopt=SGD, pwd=0.0, vwd=0.0, gamma=0.99, record=True, threads=1):
    """Runs the reference trainer. Returns a dict with logs (record) or timing."""
    pp = None if pparams is None else np.ascontiguousarray(pparams, dtype=np.float32)
    vp = None if vparams is None else np.ascontiguousarray(vparams, dtype=np.float32)
    pout = np.zeros(policy.param_count(), dtype=np.float32)
    vout = np.zeros(value.param_count() if value is not None else 1, dtype=np.float32)
    steps = C.c_longlong(0)
    vargs = value.args() if value is not None else (0, None, None, None)
    secs = lib().ref_train(algo, C.c_uint(seed), n_envs, work, iters, threads, 1 if record else 0,
                           *policy.args(), _fp(pp), popt, C.c_float(plr), C.c_float(pwd),
                           *vargs, _fp(vp), vopt, C.c_float(vlr), C.c_float(vwd),
                           C.c_float(gamma), _fp(pout), _fp(vout), C.byref(steps))
    if secs < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    res = {"seconds": secs, "env_steps": steps.value, "policy_params": pout, "value_params": vout}
    if record:
        res["steps"] = _steps()
        res["rows"] = _rows()
        ol = []
        for i in range(lib().ref_opt_log_count()):
            it, which = C.c_int(0), C.c_int(0)
            n = lib().ref_opt_log_get(i, C.byref(it), C.byref(which), None, None)
            g = np.zeros(n, dtype=np.float32)
            p = np.zeros(n, dtype=np.float32)
            lib().ref_opt_log_get(i, C.byref(it), C.byref(which), _fp(g), _fp(p))
            ol.append({"iter": it.value, "which": which.value, "grad": g, "params": p})
        res["opt_log"] = ol
    return res


def eval_argmax(seed, net, params, episodes):
    params = np.ascontiguousarray(params, dtype=np.float32)
    steps = C.c_longlong(0)
    r = lib().ref_eval_argmax(C.c_uint(seed), *net.args(), _fp(params), episodes, C.byref(steps))
    if r < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return r, steps.value
