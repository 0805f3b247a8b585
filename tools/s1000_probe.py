"""S1000 (P=1000, V=3000, highuniform, arrival 1.6): fused best-fit step time and DRL-VMP act+step rollout time."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vm-placement-migration-gym_b200"))
import numpy as np, torch
from vmgym import Config, VecVmEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
TEAM = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # warps teaming on one env (0 = auto)
DRL = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) != 0
from vmgym import _native as nv
BITS = int(sys.argv[4]) if len(sys.argv) > 4 else 7
nv.lib().vmgym_set_tuning(TEAM, BITS)
kw = dict(pms=1000, vms=3000, arrival_rate=1.6, service_length=1000, training_steps=10000, eval_steps=100000, seed=0,
          reward_function="wr", sequence="highuniform", allow_null_action=True)
vec = VecVmEnv(Config(**kw), N, rng="philox")
t0 = time.perf_counter(); vec.agent_step("bestfit", 3000, want_obs=False, want_action=False, want_valid=False); torch.cuda.synchronize()
print(f"team {TEAM} bits {BITS}: warm-up 3000 steps x {N} envs: {time.perf_counter() - t0:.2f} s")
for steps in (1, 100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    vec.agent_step("bestfit", steps, want_obs=True, want_action=False, want_valid=False)
    e0.record()
    for _ in range(5):
        vec.agent_step("bestfit", steps, want_obs=True, want_action=False, want_valid=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"fused best-fit, {steps} step(s)/launch: {ms:.3f} ms -> {N * steps / ms / 1e3:.2f} M env-steps/s")
c = vec.counters()
print("waiting per env", float(np.mean(c["slot_counts"] & 0xffff)), "empty", float(np.mean(c["slot_counts"] >> 16)))
if not DRL:
    sys.exit(0)
from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig
torch.set_float32_matmul_precision('high')      # as the reference's main.py:45
agent = DRLVMPAgent(vec, DRLVMPConfig(hidden_size=512))
agent.eval()
obs = vec.observe()
a = agent.act(obs); obs, *_ = vec.step(a, want_valid=False)      # graph capture / allocations
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2):
    a = agent.act(obs)
    obs, *_ = vec.step(a, want_valid=False)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 2
print(f"DRL-VMP act+step: {dt:.3f} s per step of {N} envs -> {N / dt:.0f} env-steps/s")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2):
    a = agent.act(obs, fused=False)
    obs, *_ = vec.step(a, want_valid=False)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 2
print(f"DRL-VMP act+step (unfused graph of torch ops): {dt:.3f} s per step -> {N / dt:.0f} env-steps/s")
