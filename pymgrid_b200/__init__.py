"""pymgrid_b200 -- B200-native batched microgrid-step engine behind pymgrid's Microgrid / envs surface."""
__version__ = "0.1.0"

from .params import BatteryParams, GensetParams, GridParams, MicrogridParams  # noqa: E402,F401
from .scenario import load_pymgrid25  # noqa: E402,F401


def __getattr__(name):
    # the engine-backed classes import torch and load the CUDA extension: resolve them lazily
    if name in ("BatchedMicrogrid", "HostIO", "HostRollout"):
        from . import engine
        return getattr(engine, name)
    if name == "Microgrid":
        from .microgrid import Microgrid
        return Microgrid
    if name in ("DiscreteMicrogridEnv", "ContinuousMicrogridEnv"):
        from . import envs
        return getattr(envs, name)
    raise AttributeError(name)
