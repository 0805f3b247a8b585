"""config/10.yml shape (P=10, V=30) at 2^20 envs: fused first-fit step time, one step and 100 steps per launch."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vm-placement-migration-gym_b200"))
import numpy as np, torch, yaml
from vmgym import Config, VecVmEnv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "10.yml")))["environment"]; cfg["reward_function"] = "wr"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
agent = sys.argv[2] if len(sys.argv) > 2 else "firstfit"
BITS = int(sys.argv[3]) if len(sys.argv) > 3 else 7           # vmgym_set_tuning use_bulk bits (16 = no double-buffered records)
from vmgym import _native as nv
nv.lib().vmgym_set_tuning(0, BITS)
vec = VecVmEnv(Config(**cfg), N, rng="philox")
vec.agent_step(agent, n_steps=3000, want_obs=False, want_action=False, want_valid=False)
for steps in (1, 100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    vec.agent_step(agent, steps, want_obs=True, want_action=False, want_valid=False)
    e0.record()
    for _ in range(10):
        vec.agent_step(agent, steps, want_obs=True, want_action=False, want_valid=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"bits {BITS} {agent} N={N} {steps} step(s)/launch: {ms:.3f} ms -> {N * steps / ms / 1e6:.3f} G env-steps/s")
