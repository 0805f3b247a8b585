#!/usr/bin/env python
"""Experiment: where does the quiet-phase single-step time go? (obs on/off, envs per launch)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import numpy as np, torch, yaml
from vmgym import Config, VecVmEnv
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def med(g, n=30, do_flush=True):
    ts = []
    for i in range(n):
        if do_flush: flush.fill_(i)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    return np.median(ts), np.min(ts)
null = torch.zeros(1, device="cuda")
gn = torch.cuda.CUDAGraph()
with torch.cuda.graph(gn):
    null.add_(1)
print("null graph (1 tiny kernel):", med(gn))
for E in (1024, 2048, 4096, 8192, 16384):
    vec = VecVmEnv(Config(**cfg), E, rng="philox")
    vec.agent_step("bestfit", n_steps=3380, want_obs=False, want_action=False, want_valid=False)   # quiet phase
    g1 = vec.capture(lambda: vec.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False))
    g0 = vec.capture(lambda: vec.agent_step("bestfit", 1, want_obs=False, want_action=False, want_valid=False))
    a, b = med(g1), med(g0)
    c = med(g1, do_flush=False)
    print(f"E={E}: with obs cold {a[0]:.1f}/{a[1]:.1f} us | no obs cold {b[0]:.1f}/{b[1]:.1f} us | with obs L2-warm {c[0]:.1f}/{c[1]:.1f} us", flush=True)
