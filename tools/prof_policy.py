#!/usr/bin/env python
"""Profiling driver for the policy-side kernels at the config/100.yml shape, 4096 envs: mask_bits_kernel, policy_fused_kernel
(streaming-sampling epilogue), heads_eval_kernel forward + backward (the PPO update's heads).  Each is launched 3 times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch, yaml
from vmgym import Config, VecVmEnv
from vmgym.ppo import FusedActorHead, PPOAgent, PPOConfig, _MaskedHeads
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
M = 4096
vec = VecVmEnv(Config(**cfg), M, rng="philox")
vec.agent_step("bestfit", n_steps=1500, want_obs=True, want_action=False, want_valid=False)
agent = PPOAgent(vec, PPOConfig(hidden_size=512, migration_ratio=0.002))
hidden = agent.model.actor[:4](vec.obs.clone()).detach()
head = FusedActorHead(agent.model.actor[4], vec.V, vec.action_dim)
logits = agent.model.actor(vec.obs.clone()).detach().requires_grad_(True)
for _ in range(3):
    bits = agent._mask_bits(0.002)
    action, lp, ent = head(hidden, bits, 1, 2)
    nlp, nent = _MaskedHeads.apply(logits, bits, action, vec._ccfg(), True)
    (nlp.sum() + nent.sum()).backward()
torch.cuda.synchronize()
print("ok")
