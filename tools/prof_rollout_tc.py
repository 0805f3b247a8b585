"""Where a PPO rollout step goes on the tensor-core path (config/100.yml shape): CUDA-event time per piece."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import torch  # noqa: E402

from bench import load_env_cfg  # noqa: E402
from vmgym import Config, VecVmEnv  # noqa: E402
from vmgym.ppo import PPOAgent, PPOConfig  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = load_env_cfg()
vec = VecVmEnv(Config(**cfg), N, rng="philox")
vec.agent_step("bestfit", n_steps=3000, want_obs=False, want_action=False, want_valid=False)
agent = PPOAgent(vec, PPOConfig(hidden_size=512, masked=True, fused_rollout=True, update_math="bf16"))
tc = agent._tc_network()
obs = vec.observe().clone()


def timed(name, fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} device {e0.elapsed_time(e1) / n * 1e3:9.1f} us   host wall {(time.perf_counter() - t0) / n * 1e6:9.1f} us")
    return out


x = timed("cast_obs", lambda: tc.cast_obs(obs, tc._buf("xr", (N, tc.Dx), torch.bfloat16)))
timed("hidden (2 gemms)", lambda: tc.hidden("actor", x, tag="r"))
h = tc.hidden("actor", x, tag="r")[1]
mask = timed("mask_bits", lambda: agent._mask_bits(-1.0))
timed("fused head", lambda: agent._fused(h, mask, 1, 1))
timed("fused_sample (all)", lambda: agent.fused_sample(obs, -1.0))
act = agent.fused_sample(obs, -1.0)[0]
timed("env step", lambda: vec.step(act, want_valid=False))
timed("torch hidden fp32", lambda: agent.model.actor[:4](obs))
