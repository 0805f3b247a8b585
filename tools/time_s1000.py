"""Fused best-fit step at the 1000-PM shape (1024 envs, one launch per step): us per step by CUDA events.
    python tools/time_s1000.py [envs] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import torch
from bench import WARM_STEPS, load_env_cfg
from vmgym import Config, VecVmEnv
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cfg = dict(load_env_cfg(), pms=1000, vms=3000, sequence="highuniform", arrival_rate=1.6)
v = VecVmEnv(Config(**cfg), E, rng="philox")
v.agent_step("bestfit", n_steps=WARM_STEPS, want_obs=False, want_action=False, want_valid=False)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(S):
        v.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / S
    print(f"{E} envs: {us:.2f} us per step = {E / us:.2f} M env-steps/s")
