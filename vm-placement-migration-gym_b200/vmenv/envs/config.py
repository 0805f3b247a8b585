from vmgym.config import Config  # noqa: F401  (reference path: vmenv/envs/config.py)
