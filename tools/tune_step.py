#!/usr/bin/env python
"""Sweep warps-per-CTA / bulk-copy for the fused step kernel (GPU box).  Prints one line per setting."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import numpy as np
import torch
import yaml
from vmgym import Config, VecVmEnv
from vmgym import _native as nv

cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]
cfg["reward_function"] = sys.argv[2] if len(sys.argv) > 2 else "wr"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
vec = VecVmEnv(Config(**cfg), E, rng="philox")
vec.agent_step("bestfit", n_steps=3000, want_obs=False, want_action=False, want_valid=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
B = 2 * (16 * 100 + 5 * 300 + 48) + 4 * 1100 + 16


def timeit(fn, iters=30, do_flush=True):
    g = vec.capture(fn)
    fn = g.replay
    for _ in range(3):
        fn()
    ts = []
    for i in range(iters):
        if do_flush:
            flush.fill_(i)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


act = vec.vm_placement.clone()
for bulk in (1,):
    for w in (0, 2, 4):
        nv.lib().vmgym_set_tuning(w, bulk)
        for name, fn in (("bestfit+step", lambda: vec.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False)),
                         ("firstfit+step", lambda: vec.agent_step("firstfit", 1, want_obs=True, want_action=False, want_valid=False)),
                         ("step(noop act)", lambda: vec.step(act, want_valid=False))):
            med, mn = timeit(fn)
            medh, mnh = timeit(fn, do_flush=False)
            print(f"E={E} bulk={bulk} warps={w} {name:16s} cold med {med*1e3:8.1f} us min {mn*1e3:8.1f} us | "
                  f"{E/med/1e3:8.2f} M env-steps/s | {B*E/med/1e6:7.1f} GB/s || L2-warm med {medh*1e3:8.1f} us "
                  f"{E/medh/1e3:8.2f} M/s", flush=True)
nv.lib().vmgym_set_tuning(0, 7)
for chunk in (10, 100):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    vec.agent_step("bestfit", chunk, want_obs=True, want_action=False, want_valid=False)
    s.record()
    for _ in range(5):
        vec.agent_step("bestfit", chunk, want_obs=True, want_action=False, want_valid=False)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"E={E} rollout chunk={chunk}: {ms:.3f} ms per launch, {E*chunk/ms/1e3:.2f} M env-steps/s", flush=True)
