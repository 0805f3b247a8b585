#!/usr/bin/env python
"""Profiling driver for the config/10.yml shape: 2^18 envs, warm-up, then single-step fused first-fit launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch, yaml
from vmgym import Config, VecVmEnv
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "10.yml")))["environment"]; cfg["reward_function"] = "wr"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
vec = VecVmEnv(Config(**cfg), N, rng="philox")
vec.agent_step("firstfit", n_steps=3000, want_obs=False, want_action=False, want_valid=False)
for i in range(3):
    vec.agent_step("firstfit", 1, want_obs=True, want_action=False, want_valid=False)
torch.cuda.synchronize()
