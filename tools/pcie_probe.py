"""PCIe probe: pinned host <-> device copy bandwidth one way, both ways at once, and the HostVecEnv loop at several group counts."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vm-placement-migration-gym_b200"))
import numpy as np
import torch
import yaml

dev = torch.device("cuda", 0)
for mb in (4.7, 19.3):
    n = int(mb * 1e6)
    h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d1, d2 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(up, down, reps=20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
            if down:
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    run(True, True, 3)
    tu, td, tb = run(True, False), run(False, True), run(True, True)
    print(f"{mb} MB: h2d {n / tu / 1e9:.1f} GB/s, d2h {n / td / 1e9:.1f} GB/s, both at once: {2 * n / tb / 1e9:.1f} GB/s total ({tb * 1e6:.0f} us)")

from vmgym import Config
from vmgym.host_vec import HostVecEnv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
E = 4096
for groups, graphs, zc, dl in ((1, True, True, True), (2, True, True, True), (4, True, True, True), (8, True, True, True), (8, True, True, False), (16, True, True, True), (4, True, False, False)):
    hv = HostVecEnv(Config(**cfg), E, groups=groups, agent="bestfit", use_graphs=graphs, zero_copy=zc, delta_obs=dl)
    hv.fast_forward(3000)
    hv.run_pipelined(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hv.run_pipelined(20)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"groups={groups} graphs={graphs} zero_copy={zc} delta_obs={dl}: {dt * 1e3:.3f} ms/step, {E / dt / 1e6:.2f} M env-steps/s, {(hv.h2d_bytes_per_step + hv.d2h_bytes_per_step) / dt / 1e9:.1f} GB/s")
    hv.close()
    del hv
# kernel times alone
from vmgym import VecVmEnv
from vmgym.agents import BestFitAgent
vec = VecVmEnv(Config(**cfg), E, rng="philox"); vec.agent_step("bestfit", 3000)
ag = BestFitAgent(vec); obs = vec.observe().clone()
for name, fn in (("agent_act", lambda: ag.act(obs)), ("step", lambda: vec.step(vec.vm_placement, want_valid=False))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us at {E} envs")
