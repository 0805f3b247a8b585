import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import torch
from bench import load_env_cfg
from vmgym import Config, VecVmEnv
from vmgym.ppo import PPOAgent, PPOConfig
N, T = 8192, 16
cfg = load_env_cfg()
vec = VecVmEnv(Config(**cfg), N, rng="philox")
vec.agent_step("bestfit", n_steps=3000, want_obs=False, want_action=False, want_valid=False)
agent = PPOAgent(vec, PPOConfig(hidden_size=512, batch_size=T, minibatch_size=T // 4, episodes=1, env_chunk=32768, masked=True, kl_max=1e9,
                                fused_rollout=True, update_math=sys.argv[1] if len(sys.argv) > 1 else "bf16"))
orig_update = agent.update
def upd(**kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig_update(**kw)
    torch.cuda.synchronize(); print(f"   update inside learn: {(time.perf_counter() - t0) * 1e3:.1f} ms")
    return r
agent.update = upd
for k in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    agent.learn(episodes=1, max_updates=1, reset=False)
    torch.cuda.synchronize(); print(f"learn call {k}: {(time.perf_counter() - t0) * 1e3:.1f} ms")
