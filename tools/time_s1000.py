"""Fused best-fit step at the 1000-PM shape: us per step by CUDA events, one launch per step and as a rotation launch
(NB phase-staggered batches of E envs, K batch steps per launch).
    python tools/time_s1000.py [envs] [steps] [batches] [K] [team warps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np, torch
from bench import WARM_STEPS, load_env_cfg
from vmgym import Config, VecVmEnv
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 4
K = int(sys.argv[4]) if len(sys.argv) > 4 else 20
if len(sys.argv) > 5:
    from vmgym import _native as nv
    nv.lib().vmgym_set_tuning(int(sys.argv[5]), 7)
cfg = dict(load_env_cfg(), pms=1000, vms=3000, sequence="highuniform", arrival_rate=1.6)
quiet = dict(want_obs=False, want_action=False, want_valid=False)
v = VecVmEnv(Config(**cfg), E, rng="philox")
v.agent_step("bestfit", n_steps=WARM_STEPS, **quiet)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(S):
        v.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / S
    print(f"per-launch  {E} envs: {us:.2f} us per step = {E / us:.2f} M env-steps/s", flush=True)
del v
# rotation: NB batches, batch b warmed to phase b * period / NB of the service period
PERIOD = int(cfg["service_length"])
v = VecVmEnv(Config(**cfg), E * NB, rng="philox", seeds=cfg["seed"] + np.arange(E * NB, dtype=np.int64))
v.agent_step("bestfit", n_steps=WARM_STEPS, **quiet)
for b in range(1, NB):
    v.agent_step("bestfit", n_steps=b * PERIOD // NB, envs=(b * E, (b + 1) * E), **quiet)
nxt = v.agent_step_rotation("bestfit", E, NB, first_batch=0)
for k in (K, K, 4 * K):
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nxt = v.agent_step_rotation("bestfit", E, k, first_batch=nxt)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / k)
    us = float(np.median(ts))
    print(f"rotation    {NB} x {E} envs, K = {k}: {us:.2f} us per batch step (min {min(ts):.2f}, max {max(ts):.2f}) = {E / us:.2f} M env-steps/s", flush=True)
