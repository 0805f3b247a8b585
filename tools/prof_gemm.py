#!/usr/bin/env python
"""Profiling driver for the tcgen05 kernels: the actor-output GEMM and the fused actor head at the config/100.yml shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch, yaml
from vmgym import Config, VecVmEnv
from vmgym.ppo import FusedActorHead, PPOAgent, PPOConfig, linear_bf16
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
M = 4096
vec = VecVmEnv(Config(**cfg), M, rng="philox")
vec.agent_step("bestfit", n_steps=1500, want_obs=True, want_action=False, want_valid=False)
agent = PPOAgent(vec, PPOConfig(hidden_size=512))
hidden = agent.model.actor[:4](vec.obs.clone()).detach()
head = FusedActorHead(agent.model.actor[4], vec.V, vec.action_dim)
w3 = agent.model.actor[4].weight.detach().bfloat16().contiguous(); b3 = agent.model.actor[4].bias.detach()
bits = agent._mask_bits(-1.0)
for _ in range(3):
    linear_bf16(hidden, w3, b3)
    head(hidden, bits, 1, 2)
torch.cuda.synchronize()
print("ok")
