"""ctypes loader + prototypes of libdfrl_b200.so (include/dfrl.h).

The library is the product: there is no Python / CPU fallback.  Importing this module fails
loudly when the shared object is missing (build it with `python __graft_entry__.py` or
`make -C dependence_free_rl_b200/csrc`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfrl_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dfrl.h")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA library is not built. Run `make -C "
        f"{os.path.join(_HERE, 'csrc')}` (nvcc, sm_100a). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

vp, i32, f32, u64, i64, sz = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_int64, C.c_size_t
pi32 = C.POINTER(C.c_int)


class EnvConfig(C.Structure):
    _fields_ = [("n_envs", i32), ("n_bins", i32), ("cap_w", i32), ("cap_h", i32),
                ("item_w", i32 * 2), ("item_h", i32 * 2), ("p_shape1", f32), ("seed", u64),
                ("env_offset", i64)]


class TrainerConfig(C.Structure):
    _fields_ = [("algo", i32), ("work", i32), ("gamma", f32), ("lambda_", f32), ("epochs", i32),
                ("kl_target", f32), ("kl_beta0", f32), ("policy_opt", i32), ("value_opt", i32),
                ("policy_lr", f32), ("value_lr", f32), ("policy_wd", f32), ("value_wd", f32),
                ("adam_beta1", f32), ("adam_beta2", f32), ("action_mode", i32), ("fused", i32)]


class TrainerStats(C.Structure):
    _fields_ = [("env_steps", C.c_longlong), ("episodes", C.c_longlong), ("reward_sum", C.c_double),
                ("last_mean_reward", C.c_double), ("kl_beta", f32)]


# name -> (restype, argtypes); every symbol include/dfrl.h declares must be listed here
PROTOTYPES = {
    "dfrl_last_error": (C.c_char_p, []),
    "dfrl_version": (C.c_char_p, []),
    "dfrl_nccl_unique_id": (i32, [vp]),
    "dfrl_init": (i32, [i32, i32, i32, vp, C.POINTER(vp)]),
    "dfrl_destroy": (i32, [vp]),
    "dfrl_sync": (i32, [vp]),
    "dfrl_stream": (vp, [vp]),
    "dfrl_device_info": (i32, [vp, pi32, pi32, pi32, C.POINTER(sz)]),
    "dfrl_malloc": (i32, [vp, sz, C.POINTER(vp)]),
    "dfrl_free": (i32, [vp, vp]),
    "dfrl_memcpy_h2d": (i32, [vp, vp, vp, sz]),
    "dfrl_memcpy_d2h": (i32, [vp, vp, vp, sz]),
    "dfrl_memcpy_d2d": (i32, [vp, vp, vp, sz]),
    "dfrl_memset": (i32, [vp, vp, i32, sz]),
    "dfrl_malloc_host": (i32, [vp, sz, C.POINTER(vp)]),
    "dfrl_free_host": (i32, [vp, vp]),
    "dfrl_profile_enable": (i32, [vp, i32]),
    "dfrl_profile_report": (i32, [vp, C.c_char_p, sz]),
    "dfrl_timer_start": (i32, [vp]),
    "dfrl_timer_stop": (i32, [vp, C.POINTER(f32)]),
    "dfrl_launch_count": (C.c_longlong, [vp]),
    "dfrl_umma_selftest": (i32, [vp, i32, i32, i32, C.POINTER(f32)]),
    "dfrl_umma_microbench": (i32, [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_longlong)]),
    "dfrl_debug_policy_clocks": (i32, [vp, vp, i32]),
    "dfrl_debug_critic_clocks": (i32, [vp, vp, i32]),
    "dfrl_debug_set_fused_ctas": (i32, [vp, i32]),
    "dfrl_debug_set_vend": (i32, [vp, i32]),
    "dfrl_trainer_fused_coverage": (i32, [vp, C.POINTER(i32)]),
    "dfrl_p2p_export": (i32, [vp, vp]),
    "dfrl_p2p_attach": (i32, [vp, vp]),
    "dfrl_p2p_attached": (i32, [vp]),
    "dfrl_env_config_default": (None, [C.POINTER(EnvConfig)]),
    "dfrl_env_create": (i32, [vp, C.POINTER(EnvConfig), C.POINTER(vp)]),
    "dfrl_env_destroy": (i32, [vp]),
    "dfrl_env_reset": (i32, [vp]),
    "dfrl_env_load_item_tape": (i32, [vp, vp, i32]),
    "dfrl_env_step": (i32, [vp, vp, vp, vp]),
    "dfrl_env_apply_one": (i32, [vp, i32, i32]),
    "dfrl_env_reset_one": (i32, [vp, i32]),
    "dfrl_env_view_one": (i32, [vp, i32, vp]),
    "dfrl_env_state_dev": (vp, [vp]),
    "dfrl_env_state_stride": (i32, [vp]),
    "dfrl_env_get_state": (i32, [vp, vp]),
    "dfrl_env_set_state": (i32, [vp, vp]),
    "dfrl_obs_encode": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "dfrl_heuristic_react": (i32, [vp, i32, vp]),
    "dfrl_heuristic_play": (i32, [vp, i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "dfrl_dense_forward": (i32, [vp, vp, i32, i32, vp, i32, vp, i32]),
    "dfrl_dense_backward": (i32, [vp, vp, i32, i32, vp, i32, vp, vp]),
    "dfrl_dense_gradient": (i32, [vp, i32, i32, vp, vp, i32, vp, i32]),
    "dfrl_relu_forward": (i32, [vp, vp, sz, vp]),
    "dfrl_relu_backward": (i32, [vp, vp, vp, sz, vp]),
    "dfrl_softmax_forward": (i32, [vp, vp, i32, i32, vp]),
    "dfrl_softmax_backward": (i32, [vp, vp, vp, i32, i32, vp]),
    "dfrl_mlp_create": (i32, [vp, i32, pi32, pi32, pi32, i32, C.POINTER(vp)]),
    "dfrl_mlp_create_shared": (i32, [vp, i32, i32, pi32, pi32, pi32, C.POINTER(vp)]),
    "dfrl_mlp_destroy": (i32, [vp]),
    "dfrl_mlp_param_count": (i32, [vp]),
    "dfrl_mlp_output_cols": (i32, [vp]),
    "dfrl_mlp_params_dev": (vp, [vp]),
    "dfrl_mlp_set_params": (i32, [vp, vp, i32]),
    "dfrl_mlp_get_params": (i32, [vp, vp, i32]),
    "dfrl_mlp_init_params": (i32, [vp, u64]),
    "dfrl_mlp_eval": (i32, [vp, vp, i32, vp]),
    "dfrl_mlp_forward_gradient": (i32, [vp, vp, i32, vp, vp, vp]),
    "dfrl_sample": (i32, [vp, vp, i32, i32, vp, vp, vp]),
    "dfrl_argmax": (i32, [vp, vp, i32, i32, vp]),
    "dfrl_returns": (i32, [vp, vp, vp, i32, i32, f32, vp, vp]),
    "dfrl_subtract_baseline": (i32, [vp, vp, vp, i32, i32, f32]),
    "dfrl_gae": (i32, [vp, vp, vp, vp, i32, i32, f32, f32, vp, vp]),
    "dfrl_loss_grad": (i32, [vp, i32, vp, vp, vp, vp, f32, i32, i32, vp]),
    "dfrl_square_loss_grad": (i32, [vp, vp, vp, i32, vp]),
    "dfrl_opt_step": (i32, [vp, i32, vp, vp, vp, i32, f32, f32, f32, f32, f32]),
    "dfrl_allreduce_sum": (i32, [vp, vp, sz]),
    "dfrl_allreduce_sum_f64": (i32, [vp, vp, sz]),
    "dfrl_barrier": (i32, [vp]),
    "dfrl_trainer_config_default": (None, [C.POINTER(TrainerConfig)]),
    "dfrl_trainer_create": (i32, [vp, C.POINTER(TrainerConfig), vp, vp, vp, C.POINTER(vp)]),
    "dfrl_trainer_destroy": (i32, [vp]),
    "dfrl_trainer_set_rates": (i32, [vp, f32, f32, f32, f32]),
    "dfrl_trainer_rollout": (i32, [vp, vp, vp, vp]),
    "dfrl_trainer_learn": (i32, [vp]),
    "dfrl_trainer_learn_phases": (i32, [vp, i32]),
    "dfrl_trainer_iterate": (i32, [vp, i32]),
    "dfrl_trainer_field_size": (i32, [vp, i32, C.POINTER(sz)]),
    "dfrl_trainer_read": (i32, [vp, i32, vp, sz]),
    "dfrl_trainer_get_stats": (i32, [vp, C.POINTER(TrainerStats)]),
    "dfrl_trainer_stats_begin": (i32, [vp]),
    "dfrl_trainer_stats_end": (i32, [vp, C.POINTER(TrainerStats)]),
    "dfrl_eval_argmax": (i32, [vp, vp, vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export the symbol
    _fn.restype = _res
    _fn.argtypes = _args


class DfrlError(RuntimeError):
    """Mirror of xeno::error (reference xeno/exception.h:12-23) for the Python binding."""


def check(rc):
    if rc != 0:
        raise DfrlError(f"dfrl error {rc}: {lib.dfrl_last_error().decode()}")
