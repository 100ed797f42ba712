"""lqr_obstacles_b200 -- B200-native per-timestep LQR-Obstacles path (CUDA + C ABI) and its host mirror.

The product is ``liblqro.so`` (csrc/, C ABI in include/lqro.h).  This package is the thin Python host
side used by the tests and the bench: ctypes bindings (api.py) and the synthetic-swarm generator
(swarm.py).  There is no CPU fallback: importing api works anywhere, creating a context needs a GPU.
"""
from .api import Lqro, LqroError, default_config, load_library  # noqa: F401
