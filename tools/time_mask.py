"""Times the mask-only kernel (PPOAgent._mask_bits) at the bench shape."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vm-placement-migration-gym_b200"))
import torch
from vmgym import Config, VecVmEnv
from vmgym.ppo import PPOAgent, PPOConfig

kw = dict(pms=100, vms=300, arrival_rate=1.8182, service_length=1000, training_steps=10000, eval_steps=100000,
          reward_function="wr", allow_null_action=True)
vec = VecVmEnv(Config(**kw), 4096, rng="philox")
vec.agent_step("bestfit", n_steps=1500)
agent = PPOAgent(vec, PPOConfig(hidden_size=256, migration_ratio=0.002))
for ratio in (-1.0, 0.002):
    for _ in range(3):
        agent._mask_bits(ratio)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        agent._mask_bits(ratio)
    e1.record(); torch.cuda.synchronize()
    print(f"mask_bits ratio={ratio}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
