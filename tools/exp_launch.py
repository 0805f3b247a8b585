#!/usr/bin/env python
"""Experiment: event-timed cost of a graph replay vs a plain launch (null kernel and the fused step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import numpy as np, torch, yaml
from vmgym import Config, VecVmEnv
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def med(fn, n=40):
    ts = []
    for i in range(n):
        flush.fill_(i)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    return f"{np.median(ts):.1f}/{np.min(ts):.1f} us"
null = torch.zeros(1, device="cuda")
gn = torch.cuda.CUDAGraph()
with torch.cuda.graph(gn):
    null.add_(1)
print("null: graph", med(gn.replay), "| plain", med(lambda: null.add_(1)), "| nothing", med(lambda: None))
vec = VecVmEnv(Config(**cfg), 4096, rng="philox")
vec.agent_step("bestfit", n_steps=3380, want_obs=False, want_action=False, want_valid=False)
fn = lambda: vec.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False)
g = vec.capture(fn)
print("step: graph", med(g.replay), "| plain", med(fn))
