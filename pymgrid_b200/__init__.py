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
    if name in ("ComposedBatch", "ComposedMicrogrid", "Composition"):      # any module list (include/pymgrid_b200_compose.h)
        from . import compose
        return getattr(compose, name)
    if name == "RuleBasedControl":
        from .algos import RuleBasedControl
        return RuleBasedControl
    if name in ("algos", "compose", "engine", "envs", "generator", "microgrid", "modules", "trajectory", "views"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
