#!/usr/bin/env python
"""Time the hand-written tcgen05 actor-output GEMM against cuBLAS (torch) on the PPO shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch
from vmgym.ppo import linear_bf16

def t(fn, it=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it

for M in (512, 4096, 8192):
    N, K = 30600, 512
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16(); b = torch.randn(N, device="cuda")
    out32 = torch.empty(M, N, device="cuda")
    ms_mine = t(lambda: linear_bf16(a, w, b))
    ms_cublas_bf16 = t(lambda: torch.addmm(b.bfloat16(), a, w.T))
    a32, w32 = a.float(), w.float()
    ms_cublas_f32out = t(lambda: torch.addmm(b, a32, w32.T, out=out32))
    fl = 2.0 * M * N * K
    print(f"M={M}: tcgen05 kernel {ms_mine:.3f} ms = {fl/ms_mine/1e9:.0f} TFLOP/s ({M*N*4/ms_mine/1e6:.0f} GB/s of fp32 output) | "
          f"cuBLAS bf16->bf16 {ms_cublas_bf16:.3f} ms = {fl/ms_cublas_bf16/1e9:.0f} TF | cuBLAS fp32(TF32?)->fp32 {ms_cublas_f32out:.3f} ms", flush=True)

# fused actor head vs GEMM + stand-alone heads kernel (config/100.yml shape: V=300, A=102, K=512)
import ctypes as C
from vmgym import Config, VecVmEnv
from vmgym import _native as nv
from vmgym.ppo import FusedActorHead, PPOAgent, PPOConfig
import yaml
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
for M in (4096,):
    vec = VecVmEnv(Config(**cfg), M, rng="philox")
    vec.agent_step("bestfit", n_steps=2000, want_obs=True, want_action=False, want_valid=False)
    agent = PPOAgent(vec, PPOConfig(hidden_size=512))
    obs = vec.obs.clone()
    hidden = agent.model.actor[:4](obs).detach()
    head = FusedActorHead(agent.model.actor[4], vec.V, vec.action_dim)
    w3 = agent.model.actor[4].weight.detach().bfloat16().contiguous(); b3 = agent.model.actor[4].bias.detach()
    bits = agent._mask_bits(-1.0)
    ms_mask = t(lambda: agent._mask_bits(-1.0))
    ms_fused = t(lambda: head(hidden, bits, 1, 2))
    def unfused():
        lg = linear_bf16(hidden, w3, b3)
        return agent._heads(lg, -1.0, want_mask=False)
    ms_unfused = t(unfused)
    ms_hidden = t(lambda: agent.model.actor[:4](obs))
    ms_step = t(lambda: vec.step(vec.agent_action, want_valid=False))
    fl = 2.0 * M * 300 * 128 * 512
    print(f"M={M}: fused head {ms_fused:.3f} ms ({fl/ms_fused/1e9:.0f} TFLOP/s incl. padding) | GEMM+heads kernel {ms_unfused:.3f} ms | "
          f"mask bits {ms_mask:.3f} ms | hidden layers (torch) {ms_hidden:.3f} ms | env step {ms_step:.3f} ms", flush=True)
