#!/usr/bin/env python
"""Fused best-fit step kernel: time per launch and roofline fraction vs envs per launch (config/100.yml, Philox), one launch
per step with L2 flushed before each timed launch, 10 timed launches spread over a service period."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import numpy as np, torch, yaml
from vmgym import Config, VecVmEnv
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6553.3) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6553.3
B = 2 * (16 * 100 + 5 * 300 + 48) + 4 * (3 * 300 + 2 * 100) + 16
from vmgym import _native as nv
BITS = int(sys.argv[1]) if len(sys.argv) > 1 else 7          # vmgym_set_tuning use_bulk bits (16 = no double-buffered records)
SIZES = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072]
nv.lib().vmgym_set_tuning(0, BITS)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("tuning bits", BITS)
print("| envs per launch | us per launch | M env-steps/s | algorithmic GB/s | fraction of %.0f GB/s |" % peak)
print("|---|---|---|---|---|")
for E in SIZES:
    vec = VecVmEnv(Config(**cfg), E, rng="philox")
    vec.agent_step("bestfit", n_steps=3000, want_obs=False, want_action=False, want_valid=False)
    g = vec.capture(lambda: vec.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False))
    ts = []
    for k in range(10):
        vec.agent_step("bestfit", 99, want_obs=False, want_action=False, want_valid=False)
        flush.fill_(k)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); g.replay(); s1.record(); torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    us = float(np.mean(ts)) * 1e3
    gbs = B * E / (us * 1e-6) / 1e9
    print(f"| {E} | {us:.1f} | {E / us:.1f} | {gbs:.0f} | {gbs / peak:.3f} |")
    del vec, g
