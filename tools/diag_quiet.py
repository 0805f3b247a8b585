#!/usr/bin/env python
"""Diagnostic: per-step fraction of QUIET envs / changed envs at saturation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import numpy as np, torch, yaml
from vmgym import Config, VecVmEnv
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]
cfg["reward_function"] = "wr"
E = 1024
vec = VecVmEnv(Config(**cfg), E, rng="philox")
vec.agent_step("bestfit", n_steps=3000, want_obs=False, want_action=False, want_valid=False)
prev = vec.counters()
for t in range(4):
    vec.agent_step("bestfit", 1, want_obs=False, want_action=False, want_valid=False)
    c = vec.counters()
    st = c["status"]
    changed = (c["served_requests"] != prev["served_requests"]) | (c["place_actions"] != prev["place_actions"]) | (c["admission_pos"] != prev["admission_pos"])
    sl = c["slot_counts"]
    print({k: float((c[k] != prev[k]).mean()) for k in ("served_requests", "place_actions", "admission_pos", "dropped_requests", "total_requests")},
          "served/step", float((c["served_requests"] - prev["served_requests"]).mean()))
    print(f"t={t} quiet={((st & 2) != 0).mean():.3f} key0={(((st & 2) != 0) & ((st & 0xff00) == 0)).mean():.3f} changed={changed.mean():.3f} "
          f"n_waiting={np.mean(sl & 0xffff):.1f} n_empty={np.mean(sl >> 16):.2f} rejected_field={np.mean((st >> 16) & 0xffff):.3f}")
    prev = c
