"""B200-native VSS (3v3 robot soccer) hot path: fused CUDA env step, agent views, GAE, PPO.

Drop-in for the `envs/vss.py` task, the `envs/wrappers.py` views and the
`ppo_continuous_action_isaacgym.py` loop of FelipeMartins96/rsoccer-isaac-cleanrl, with
IsaacGym/PhysX and the torch-jit reward/obs code replaced by `libvss_b200.so` (sm_100a).
There is no CPU fallback: every compute entry point needs the built library and a B200.
"""
from . import _lib  # noqa: F401
from ._lib import VssParams, default_params, load_library  # noqa: F401
from .engine import Engine  # noqa: F401

__all__ = ["Engine", "VssParams", "default_params", "load_library"]
