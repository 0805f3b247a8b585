"""vmgym — B200-native batched VM placement/migration env (drop-in for the reference's vmenv hot path)."""
from .config import Config  # noqa: F401


def __getattr__(name):
    # lazy: importing the package must not require CUDA (the CPU suite only checks the ABI)
    if name == "VecVmEnv":
        from .vec_env import VecVmEnv
        return VecVmEnv
    if name == "VmEnv":
        from .env import VmEnv
        return VmEnv
    if name in ("FirstFitAgent", "BestFitAgent"):
        from . import agents
        return getattr(agents, name)
    raise AttributeError(name)
