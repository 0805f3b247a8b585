#!/usr/bin/env python
"""Profiling driver: warm a 4096-env batch to saturation, then issue a few single-step fused launches (L2 flushed)
and one 100-step launch.  Run under ncu with `-k regex:step_kernel -s 1` to skip the warm-up launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch
import yaml
from vmgym import Config, VecVmEnv

cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]
cfg["reward_function"] = "wr"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
agent = sys.argv[2] if len(sys.argv) > 2 else "bestfit"
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3000     # 3000: departure-wave phase, 3500: quiet phase
vec = VecVmEnv(Config(**cfg), E, rng="philox")
vec.agent_step(agent, n_steps=warm, want_obs=False, want_action=False, want_valid=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(3):
    flush.fill_(i)
    vec.agent_step(agent, 1, want_obs=True, want_action=False, want_valid=False)
vec.agent_step(agent, 100, want_obs=True, want_action=False, want_valid=False)
torch.cuda.synchronize()
st = vec._scalars_i32[:, 8].cpu().numpy()
print("quiet fraction", float(((st & 2) != 0).mean()), "key0 fraction", float((((st & 2) != 0) & ((st & 0xff00) == 0)).mean()))
