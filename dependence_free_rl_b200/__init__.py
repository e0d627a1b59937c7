"""dependence_free_rl_b200 -- B200 (sm_100a) implementation of the on-policy bin-packing training
loop of beehover/dependence_free_rl, behind the C ABI of include/dfrl.h.

Importing the package loads dependence_free_rl_b200/libdfrl_b200.so and fails loudly when it is
missing; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is not built)
from .api import *  # noqa: F401,F403
from .api import Context, DeviceArray, Environment, Model, Trainer, eval_argmax, fc_layers, conv_layers  # noqa: F401
