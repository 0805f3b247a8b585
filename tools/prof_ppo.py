"""Kernel-time breakdown of PPOAgent.learn (rollout + update) at the bench's ppo_train shape, via torch.profiler."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vm-placement-migration-gym_b200"))
import numpy as np, torch, yaml
from torch.profiler import profile, ProfilerActivity
from vmgym import Config, VecVmEnv
from vmgym.ppo import PPOAgent, PPOConfig
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
Np = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Tp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
vp = VecVmEnv(Config(**cfg), Np, rng="philox")
torch.set_float32_matmul_precision("high")
agent = PPOAgent(vp, PPOConfig(hidden_size=512, batch_size=Tp, minibatch_size=Tp // 4, episodes=1, env_chunk=chunk, masked=True, kl_max=1e9,
                               fused_rollout=True))
agent.learn(episodes=1, max_updates=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
agent.learn(episodes=1, max_updates=2)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"N={Np} T={Tp} chunk={chunk}: {dt:.3f} s for 2 updates -> {2 * Np * Tp / dt:.0f} env-steps/s")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    agent.learn(episodes=1, max_updates=1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
