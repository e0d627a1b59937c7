"""Thin object layer over the C ABI (include/dfrl.h) for tests and bench.py.

Names follow the reference's vocabulary: Environment (bp::environment), Model (xylo::model),
Trainer (the learner + replay buffer + trainer-main loop).  numpy arrays go in and out; all
compute happens in libdfrl_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check

DENSE, CONV1D, RELU, SOFTMAX, SOFTMAX_CE = 0, 1, 2, 3, 4
REINFORCE, ACTOR_CRITIC, PPO, KL_PPO = 0, 1, 2, 3
SGD, MOMENTUM, ADAM = 0, 1, 2
LOSS_SOFTMAX_LOG, LOSS_CLIPPED, LOSS_KL = 0, 1, 2
ACT_SAMPLE, ACT_ARGMAX, ACT_FORCED = 0, 1, 2
HEUR_RANDOM, HEUR_FIRSTFIT, HEUR_BESTFIT, HEUR_MINWASTE = 0, 1, 2, 3
(F_REC_STATE, F_REC_ACTION, F_REC_DONE, F_REC_PROBS, F_REC_LEN, F_ADVANTAGE, F_VALUE_TARGET,
 F_POLICY_GRAD, F_VALUE_GRAD, F_POLICY_GRAD_LOG, F_OBS_START) = range(11)
PHASE_VALUE, PHASE_ADVANTAGE, PHASE_POLICY, PHASE_ALL = 1, 2, 4, 7
FUSED_ROLLOUT, FUSED_CRITIC, FUSED_POLICY, FUSED_GRAPH = 1, 2, 4, 8


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    def __init__(self, device=0, nranks=1, rank=0, nccl_id=None):
        h = C.c_void_p()
        idbuf = None
        if nccl_id is not None:
            idbuf = (C.c_char * 128).from_buffer_copy(nccl_id)
        check(lib.dfrl_init(device, nranks, rank, idbuf, C.byref(h)))
        self.h = h
        self.nranks, self.rank = nranks, rank
        self._bufs = []

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        check(lib.dfrl_nccl_unique_id(buf))
        return bytes(buf)

    # ---- flat-gradient exchange over NVLink peer memory (dfrl_p2p_*) ----
    def p2p_export(self):
        """64-byte CUDA IPC handle of this rank's exchange buffer (gather these over all ranks)."""
        buf = (C.c_char * 64)()
        check(lib.dfrl_p2p_export(self.h, buf))
        return bytes(buf)

    def p2p_attach(self, handles):
        """handles: the exported handle of every rank, in rank order."""
        assert len(handles) == self.nranks
        buf = (C.c_char * (64 * self.nranks)).from_buffer_copy(b"".join(handles))
        check(lib.dfrl_p2p_attach(self.h, buf))

    def p2p_attached(self):
        return bool(lib.dfrl_p2p_attached(self.h))

    def close(self):
        if self.h:
            lib.dfrl_destroy(self.h)
            self.h = None

    def sync(self):
        check(lib.dfrl_sync(self.h))

    def device_info(self):
        sm, ma, mi, hb = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        check(lib.dfrl_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(hb)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hb.value}

    def launches(self):
        return lib.dfrl_launch_count(self.h)

    # ---- device memory ----
    def malloc(self, nbytes):
        p = C.c_void_p()
        check(lib.dfrl_malloc(self.h, nbytes, C.byref(p)))
        return p

    def free(self, p):
        check(lib.dfrl_free(self.h, p))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(max(arr.nbytes, 1))
        check(lib.dfrl_memcpy_h2d(self.h, p, _ptr(arr), arr.nbytes))
        return DeviceArray(self, p, arr.shape, arr.dtype)

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        return DeviceArray(self, self.malloc(max(n, 1)), tuple(shape), dtype)

    def zeros(self, shape, dtype):
        a = self.empty(shape, dtype)
        check(lib.dfrl_memset(self.h, a.p, 0, a.nbytes))
        return a

    def timer_start(self):
        check(lib.dfrl_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib.dfrl_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def allreduce_sum(self, darr):
        check(lib.dfrl_allreduce_sum(self.h, darr.p, darr.size))

    def barrier(self):
        check(lib.dfrl_barrier(self.h))


class DeviceArray:
    """memory_blob(on_device = true) (reference tensor.cc:78-102) with a shape."""

    def __init__(self, ctx, p, shape, dtype):
        self.ctx, self.p, self.shape, self.dtype = ctx, p, tuple(shape), np.dtype(dtype)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    def get(self):
        out = np.empty(self.shape, self.dtype)
        check(lib.dfrl_memcpy_d2h(self.ctx.h, _ptr(out), self.p, self.nbytes))
        return out

    def set(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        assert arr.size == self.size
        check(lib.dfrl_memcpy_h2d(self.ctx.h, self.p, _ptr(arr), arr.nbytes))

    def free(self):
        if self.p:
            self.ctx.free(self.p)
            self.p = None


class Environment:
    """Batched bp::environment (+ bp::agent's game_over / reward / reset bookkeeping)."""

    def __init__(self, ctx, n_envs, n_bins=8, seed=1234, env_offset=0, **kw):
        cfg = _lib.EnvConfig()
        lib.dfrl_env_config_default(C.byref(cfg))
        cfg.n_envs, cfg.n_bins, cfg.seed, cfg.env_offset = n_envs, n_bins, seed, env_offset
        for k, v in kw.items():
            if k in ("item_w", "item_h"):
                getattr(cfg, k)[0], getattr(cfg, k)[1] = v
            else:
                setattr(cfg, k, v)
        h = C.c_void_p()
        check(lib.dfrl_env_create(ctx.h, C.byref(cfg), C.byref(h)))
        self.ctx, self.h, self.n, self.B, self.cfg = ctx, h, n_envs, n_bins, cfg
        self.stride = lib.dfrl_env_state_stride(h)

    def close(self):
        if self.h:
            lib.dfrl_env_destroy(self.h)
            self.h = None

    def reset(self):
        check(lib.dfrl_env_reset(self.h))

    def load_item_tape(self, tape):
        if tape is None:
            check(lib.dfrl_env_load_item_tape(self.h, None, 0))
            return
        tape = np.ascontiguousarray(tape, dtype=np.uint8)
        assert tape.shape[0] == self.n
        check(lib.dfrl_env_load_item_tape(self.h, _ptr(tape), tape.shape[1]))

    def state(self):
        st = np.empty((2 * self.B + 2, self.n), np.int8)
        check(lib.dfrl_env_get_state(self.h, _ptr(st)))
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.int8)
        assert st.shape == (2 * self.B + 2, self.n)
        check(lib.dfrl_env_set_state(self.h, _ptr(st)))

    def step(self, actions, want_terminal=False):
        """actions: uint8 [N] (numpy or DeviceArray). Returns done [N] (and terminal planes)."""
        a = actions if isinstance(actions, DeviceArray) else self.ctx.to_device(np.asarray(actions, np.uint8))
        done = self.ctx.empty((self.n,), np.uint8)
        term = self.ctx.empty((2 * self.B + 2, self.stride), np.int8) if want_terminal else None
        check(lib.dfrl_env_step(self.h, a.p, done.p, term.p if term else None))
        d = done.get()
        t = term.get()[:, :self.n] if term else None
        done.free()
        if term:
            term.free()
        if not isinstance(actions, DeviceArray):
            a.free()
        return (d, t) if want_terminal else d

    def obs(self):
        o = self.ctx.empty((self.n, 4 * self.B), np.float32)
        check(lib.dfrl_obs_encode(self.ctx.h, lib.dfrl_env_state_dev(self.h), self.n, self.stride,
                                  self.B, self.cfg.cap_w, self.cfg.cap_h, o.p))
        out = o.get()
        o.free()
        return out

    def heuristic_react(self, kind):
        a = self.ctx.empty((self.n,), np.uint8)
        check(lib.dfrl_heuristic_react(self.h, kind, a.p))
        out = a.get()
        a.free()
        return out

    def heuristic_play(self, kind, episodes):
        tot, steps = C.c_double(), C.c_longlong()
        check(lib.dfrl_heuristic_play(self.h, kind, episodes, C.byref(tot), C.byref(steps)))
        return tot.value, steps.value


class Model:
    """xylo::model with device parameters. layers: [(kind, in, out), ...]."""

    def __init__(self, ctx, layers, input_cols):
        self.layers = [tuple(int(v) for v in l) for l in layers]
        k = np.array([l[0] for l in self.layers], np.int32)
        i = np.array([l[1] for l in self.layers], np.int32)
        o = np.array([l[2] for l in self.layers], np.int32)
        h = C.c_void_p()
        check(lib.dfrl_mlp_create(ctx.h, len(self.layers), k.ctypes.data_as(_lib.pi32),
                                  i.ctypes.data_as(_lib.pi32), o.ctypes.data_as(_lib.pi32),
                                  input_cols, C.byref(h)))
        self.ctx, self.h, self.input_cols = ctx, h, input_cols
        self.output_cols = lib.dfrl_mlp_output_cols(h)

    @classmethod
    def shared(cls, trunk, n_shared, head_layers):
        """A model whose first n_shared layers are `trunk`'s (same device parameters) followed by its
        own head_layers (dfrl_mlp_create_shared): the shared-trunk policy / value pair of BASELINE
        configs[2]. Both models then report the family's whole flat parameter vector."""
        self = cls.__new__(cls)
        head = [tuple(int(v) for v in l) for l in head_layers]
        self.layers = trunk.layers[:n_shared] + head
        k = np.array([l[0] for l in head], np.int32)
        i = np.array([l[1] for l in head], np.int32)
        o = np.array([l[2] for l in head], np.int32)
        h = C.c_void_p()
        check(lib.dfrl_mlp_create_shared(trunk.h, n_shared, len(head), k.ctypes.data_as(_lib.pi32),
                                         i.ctypes.data_as(_lib.pi32), o.ctypes.data_as(_lib.pi32), C.byref(h)))
        self.ctx, self.h, self.input_cols = trunk.ctx, h, trunk.input_cols
        self.output_cols = lib.dfrl_mlp_output_cols(h)
        self._family = trunk
        trunk._sharers = getattr(trunk, "_sharers", []) + [self]
        return self

    @property
    def n_params(self):
        return lib.dfrl_mlp_param_count(self.h)

    def close(self):
        if self.h:
            check(lib.dfrl_mlp_destroy(self.h))
            self.h = None

    def set_parameters(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32)
        check(lib.dfrl_mlp_set_params(self.h, _ptr(p), p.size))

    def parameters(self):
        p = np.empty(self.n_params, np.float32)
        check(lib.dfrl_mlp_get_params(self.h, _ptr(p), p.size))
        return p

    def init_parameters(self, seed):
        check(lib.dfrl_mlp_init_params(self.h, seed))

    def eval(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        dx = self.ctx.to_device(x)
        dy = self.ctx.empty((x.shape[0], self.output_cols), np.float32)
        check(lib.dfrl_mlp_eval(self.h, dx.p, x.shape[0], dy.p))
        out = dy.get()
        dx.free()
        dy.free()
        return out

    def forward_gradient(self, x, dy):
        x = np.ascontiguousarray(x, dtype=np.float32)
        dy = np.ascontiguousarray(dy, dtype=np.float32)
        dx_, ddy = self.ctx.to_device(x), self.ctx.to_device(dy)
        g = self.ctx.empty((self.n_params,), np.float32)
        o = self.ctx.empty((x.shape[0], self.output_cols), np.float32)
        check(lib.dfrl_mlp_forward_gradient(self.h, dx_.p, x.shape[0], ddy.p, g.p, o.p))
        res = g.get(), o.get()
        for a in (dx_, ddy, g, o):
            a.free()
        return res


def fc_layers(dims, last=None):
    layers = []
    for i in range(len(dims) - 1):
        layers.append((DENSE, dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append((RELU, 0, 0))
    if last is not None:
        layers.append((last, 0, 0))
    return layers


def conv_layers(chans, last=None):
    layers = []
    for i in range(len(chans) - 1):
        layers.append((CONV1D, chans[i], chans[i + 1]))
        if i < len(chans) - 2:
            layers.append((RELU, 0, 0))
    if last is not None:
        layers.append((last, 0, 0))
    return layers


class Trainer:
    """Replay buffer + learner + the trainer-main loop (rollout; learn; forget)."""

    def __init__(self, ctx, env, policy, value=None, algo=PPO, work=4, **kw):
        cfg = _lib.TrainerConfig()
        lib.dfrl_trainer_config_default(C.byref(cfg))
        cfg.algo, cfg.work = algo, work
        for k, v in kw.items():
            if k == "lambda":
                k = "lambda_"
            if not hasattr(cfg, k):
                raise TypeError(f"unknown trainer option {k}")
            setattr(cfg, k, v)
        h = C.c_void_p()
        check(lib.dfrl_trainer_create(ctx.h, C.byref(cfg), env.h, policy.h,
                                      value.h if value is not None else None, C.byref(h)))
        self.ctx, self.h, self.env, self.policy, self.value, self.cfg = ctx, h, env, policy, value, cfg

    def close(self):
        if self.h:
            lib.dfrl_trainer_destroy(self.h)
            self.h = None

    def rollout(self, items=None, actions=None, u=None):
        it = None if items is None else np.ascontiguousarray(items, dtype=np.uint8)
        ac = None if actions is None else np.ascontiguousarray(actions, dtype=np.uint8)
        uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        check(lib.dfrl_trainer_rollout(self.h, _ptr(it), _ptr(ac), _ptr(uu)))

    def rollout_raw(self, items_ptr, actions_ptr, u_ptr):
        check(lib.dfrl_trainer_rollout(self.h, items_ptr, actions_ptr, u_ptr))

    def learn(self, phases=None):
        if phases is None:
            check(lib.dfrl_trainer_learn(self.h))
        else:
            check(lib.dfrl_trainer_learn_phases(self.h, phases))

    def iterate(self, iters):
        check(lib.dfrl_trainer_iterate(self.h, iters))

    def set_rates(self, policy_lr, value_lr=0.0, policy_wd=0.0, value_wd=0.0):
        """optimizer::set_rate (nn.h:592) on the live learner; optimizer state is kept."""
        check(lib.dfrl_trainer_set_rates(self.h, policy_lr, policy_wd, value_lr, value_wd))

    def read(self, field):
        nb = C.c_size_t()
        check(lib.dfrl_trainer_field_size(self.h, field, C.byref(nb)))
        buf = np.empty(nb.value, np.uint8)
        check(lib.dfrl_trainer_read(self.h, field, _ptr(buf), nb.value))
        n, B = self.env.n, self.env.B
        if field == F_REC_STATE:
            return buf.view(np.int8).reshape(-1, 2 * B + 2, n)
        if field in (F_REC_ACTION, F_REC_DONE):
            return buf.reshape(-1, n)
        if field == F_REC_PROBS:
            return buf.view(np.float32).reshape(-1, n, B)
        if field == F_REC_LEN:
            return buf.view(np.int32)
        if field in (F_ADVANTAGE, F_VALUE_TARGET):
            return buf.view(np.float32).reshape(-1, n)
        if field == F_POLICY_GRAD_LOG:
            return buf.view(np.float32).reshape(-1, self.policy.n_params)
        if field == F_OBS_START:
            return buf.view(np.float32).reshape(-1, n, 4 * B)
        return buf.view(np.float32)

    @staticmethod
    def _stats_dict(s):
        return {"env_steps": s.env_steps, "episodes": s.episodes, "reward_sum": s.reward_sum,
                "last_mean_reward": s.last_mean_reward, "kl_beta": s.kl_beta}

    def fused_coverage(self):
        """Mask of FUSED_ROLLOUT | FUSED_CRITIC | FUSED_POLICY | FUSED_GRAPH: phases on the tcgen05 kernels."""
        m = C.c_int()
        check(lib.dfrl_trainer_fused_coverage(self.h, C.byref(m)))
        return m.value

    def fused_covers_iteration(self):
        return (self.fused_coverage() & 7) == 7

    def stats(self):
        s = _lib.TrainerStats()
        check(lib.dfrl_trainer_get_stats(self.h, C.byref(s)))
        return self._stats_dict(s)

    def stats_begin(self):
        """Enqueue the device->host read of the counters (non-blocking); pair with stats_end()."""
        check(lib.dfrl_trainer_stats_begin(self.h))

    def stats_end(self):
        """Wait for the oldest stats_begin() and return its result."""
        s = _lib.TrainerStats()
        check(lib.dfrl_trainer_stats_end(self.h, C.byref(s)))
        return self._stats_dict(s)


def eval_argmax(ctx, env, policy, episodes):
    mean, steps = C.c_double(), C.c_longlong()
    check(lib.dfrl_eval_argmax(ctx.h, env.h, policy.h, episodes, C.byref(mean), C.byref(steps)))
    return mean.value, steps.value
