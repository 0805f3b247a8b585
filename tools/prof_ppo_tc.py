"""Timing breakdown of one PPO update on the tensor-core path (config/100.yml shape).
    python tools/prof_ppo_tc.py [envs] [T] [chunk]
Prints per-kernel-class CUDA time from torch.profiler plus the end-to-end update time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import load_env_cfg  # noqa: E402
from vmgym import Config, VecVmEnv  # noqa: E402
from vmgym.ppo import PPOAgent, PPOConfig  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
mode = sys.argv[4] if len(sys.argv) > 4 else "bf16"
cfg = load_env_cfg()
vec = VecVmEnv(Config(**cfg), N, rng="philox")
vec.agent_step("bestfit", n_steps=3000, want_obs=False, want_action=False, want_valid=False)
torch.set_float32_matmul_precision("high")
agent = PPOAgent(vec, PPOConfig(hidden_size=512, batch_size=T, minibatch_size=max(1, T // 4), episodes=1, env_chunk=chunk, masked=True,
                                kl_max=1e9, fused_rollout=True, update_math=mode))
agent.learn(episodes=1, max_updates=1, reset=False)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
agent.learn(episodes=1, max_updates=1, reset=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"rollout + update: {dt * 1e3:.1f} ms for {N * T} env-steps = {N * T / dt / 1e6:.3f} M env-steps/s ({mode})")
buf = agent._rollout
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); agent.update(**buf); e1.record(); torch.cuda.synchronize()
print(f"update alone: {e0.elapsed_time(e1):.1f} ms")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    agent.update(**buf)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:14]
tot = sum(r.device_time_total for r in prof.key_averages())
for r in rows:
    print(f"{r.device_time_total / 1e3:9.2f} ms {100 * r.device_time_total / tot:5.1f}%  x{r.count:<5d} {r.key[:110]}")
