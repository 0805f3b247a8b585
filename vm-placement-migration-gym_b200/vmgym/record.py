"""Record — what Base.test returns in the reference (src/record.py).  The reference keeps every step's info in Python
lists (V x T placements, T x P utilisation) and derives the summary afterwards; here the same quantities are running
sums / histograms accumulated by the kernels (VecVmEnv.summary), so a Record is the summary plus the configs."""
from __future__ import annotations

import dataclasses
import json
import os

import numpy as np

SUMMARY_KEYS = ["total rewards", "total served VMs", "total requests", "total cpu requested", "total memory requested",
                "total suspend actions", "total place actions", "average VM life", "average pending", "median pending",
                "max pending", "average slowdown", "median slowdown", "max slowdown", "drop rate", "cpu mean", "cpu mean target",
                "cpu std", "memory mean", "memory mean target", "memory std", "rank mean"]        # record.py:110-134, in order


def _asdict(cfg):
    if cfg is None:
        return None
    return dataclasses.asdict(cfg) if dataclasses.is_dataclass(cfg) else dict(cfg)


class Record:
    def __init__(self, agent: str, env_config, agent_config, summary: dict):
        self.agent = agent
        self.env_config = _asdict(env_config)
        self.agent_config = _asdict(agent_config)
        self.raw = {k: np.asarray(v) for k, v in summary.items()}            # unrounded, one entry per env of the batch
        self.num_envs = int(next(iter(self.raw.values())).shape[0]) if self.raw else 0

    def get_summary(self, env: int = 0):
        """record.py:110-134 for one env of the batch: the same keys, rounded to 3 decimals like the reference."""
        out = {}
        for k in SUMMARY_KEYS:
            if k in self.raw:
                v = self.raw[k][env]
                is_int = k in ("total served VMs", "total requests", "total suspend actions", "total place actions")
                out[k] = int(v) if is_int else float(np.round(v, 3))
        return out

    def save(self, path: str, env: int = 0):
        """record.py:136-141 — JSON with the configs and the summary (the per-step lists of the reference are not kept)."""
        d = os.path.dirname(os.path.abspath(path))
        os.makedirs(d, exist_ok=True)
        with open(path, "w") as f:
            json.dump(dict(agent=self.agent, env_config=self.env_config, agent_config=self.agent_config,
                           summary=self.get_summary(env)), f)
