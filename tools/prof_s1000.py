#!/usr/bin/env python
"""Profiling driver for the synthetic 1000-PM shape: warm 1024 envs to saturation, then two single-step fused best-fit
launches.  Run under ncu with `-k regex:step_kernel -s 1` to skip the warm-up launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch
from vmgym import Config, VecVmEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
kw = dict(pms=1000, vms=3000, arrival_rate=1.6, service_length=1000, training_steps=10000, eval_steps=100000, seed=0,
          reward_function="wr", sequence="highuniform", allow_null_action=True)
vec = VecVmEnv(Config(**kw), N, rng="philox")
vec.agent_step("bestfit", warm, want_obs=False, want_action=False, want_valid=False)
for i in range(3):
    vec.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False)
torch.cuda.synchronize()
