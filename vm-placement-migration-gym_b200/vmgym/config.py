"""Environment configuration — the reference's 12-field dataclass (vmenv/envs/config.py:3-16), same names and
defaults, so `Config(**yaml["environment"])` keeps working (main.py:47)."""
from __future__ import annotations

from dataclasses import dataclass

REWARD_FUNCTIONS = ("wr", "ut", "kl")                 # main.py:94; "reward 1/2/3" of the paper
# env.py:211-219: uniform size range per `sequence`, in hundredths (sizes are rounded to 2 decimals)
SEQUENCE_RANGES = {"uniform": (0.1, 1.0), "lowuniform": (0.1, 0.65), "highuniform": (0.25, 1.0)}
SEQUENCE_CODES = {"uniform": (10, 100), "lowuniform": (10, 65), "highuniform": (25, 100)}


@dataclass
class Config(object):
    arrival_rate: float = 0.182      # 100% load = pms / E[size] / service_length
    service_length: float = 100
    pms: int = 10
    vms: int = 30
    training_steps: int = 500
    eval_steps: int = 100000
    seed: int = 0
    reward_function: str = "wr"
    sequence: str = "uniform"
    cap_target_util: bool = True
    beta: float = 0.5
    allow_null_action: bool = False  # the masked PPO needs the extra "stay empty" action

    def validate(self):
        if self.reward_function not in REWARD_FUNCTIONS:
            # the reference asserts at step time (env.py:155-156)
            raise AssertionError(f"Function does not exist: {self.reward_function}")
        if self.sequence not in SEQUENCE_RANGES:
            raise ValueError(f"unknown sequence {self.sequence!r}")
        if self.pms < 1 or self.vms < 1:
            raise ValueError("pms and vms must be positive")
        return self

    @property
    def action_dim(self) -> int:                      # env.py:26
        return self.pms + 2 if self.allow_null_action else self.pms + 1

    @property
    def obs_dim(self) -> int:                         # env.py:27
        return self.vms * 3 + self.pms * 2
