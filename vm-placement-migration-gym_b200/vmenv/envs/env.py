from vmgym.config import Config  # noqa: F401
from vmgym.env import VmEnv  # noqa: F401  (reference path: vmenv/envs/env.py)
from vmgym.vec_env import VecVmEnv  # noqa: F401
