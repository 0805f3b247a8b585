"""Time of ONE rotation launch of K batch steps vs K (config/100.yml, 20 phase-staggered batches of 4096 envs, fused best-fit):
median of 15 launches per K, CUDA events around each launch.   python tools/rot_k_sweep.py [K,K,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np, torch
from bench import PERIOD, WARM_STEPS, load_env_cfg
from vmgym import Config, VecVmEnv
KS = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 3, 5, 8, 12, 20, 35, 50, 100, 200]
E, NB = 4096, 20
cfg = load_env_cfg()
seeds = np.concatenate([cfg["seed"] + b * E + np.arange(E, dtype=np.int64) for b in range(NB)])
vec = VecVmEnv(Config(**cfg), NB * E, rng="philox", seeds=seeds)
for b in range(NB):
    vec.agent_step("bestfit", n_steps=WARM_STEPS + (b * PERIOD) // NB, want_obs=False, want_action=False, want_valid=False, envs=(b * E, (b + 1) * E))
nxt = vec.agent_step_rotation("bestfit", E, NB, first_batch=0)
for K in KS:
    ts = []
    for r in range(15):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nxt = vec.agent_step_rotation("bestfit", E, K, first_batch=nxt)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"K = {K:4d}: {np.median(ts):9.1f} us per launch = {np.median(ts) / K:6.2f} us per batch step (min {min(ts) / K:.2f})", flush=True)
