#!/usr/bin/env python
"""A/B harness: wave-phase cost of the fused kernel right after the 3000-step warm-up (the expensive phase):
100-step launches and single-step launches, then the same in the quiet phase 500 steps later."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import numpy as np, torch, yaml
from vmgym import Config, VecVmEnv
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]
cfg["reward_function"] = "wr"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
agent = sys.argv[2] if len(sys.argv) > 2 else "bestfit"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def measure(vec, tag):
    g = vec.capture(lambda: vec.agent_step(agent, 1, want_obs=True, want_action=False, want_valid=False))
    ts = []
    for i in range(20):
        flush.fill_(i)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    vec.agent_step(agent, 100, want_obs=True, want_action=False, want_valid=False)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"{tag}: single-step cold med {np.median(ts):.1f} us (min {np.min(ts):.1f}) = {E/np.median(ts):.1f} M/s | "
          f"100-step launch {ms:.3f} ms = {E*100/ms/1e3:.1f} M env-steps/s", flush=True)


vec = VecVmEnv(Config(**cfg), E, rng="philox")
vec.agent_step(agent, n_steps=3000, want_obs=False, want_action=False, want_valid=False)
measure(vec, f"E={E} {agent} wave ")
vec.agent_step(agent, n_steps=380, want_obs=False, want_action=False, want_valid=False)
measure(vec, f"E={E} {agent} quiet")
