"""Cycles per section of the step kernel inside the benchmark's rotation launch (needs a library built with
VMGYM_NVCC_EXTRA=-DVMGYM_PROF; warp 0 of every CTA reports).
    python tools/prof_sections_rot.py [envs_per_batch] [batches] [K] [s100|s1000]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import PERIOD, WARM_STEPS, load_env_cfg  # noqa: E402
from vmgym import Config, VecVmEnv  # noqa: E402
from vmgym import _native as nv  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 20
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
cfg = load_env_cfg()
SHAPE = sys.argv[4] if len(sys.argv) > 4 else "s100"
if SHAPE == "s1000":
    cfg = dict(cfg, pms=1000, vms=3000, sequence="highuniform", arrival_rate=1.6)
seeds = np.concatenate([cfg["seed"] + b * E + np.arange(E, dtype=np.int64) for b in range(NB)])
vec = VecVmEnv(Config(**cfg), NB * E, rng="philox", seeds=seeds)
for b in range(NB):
    vec.agent_step("bestfit", n_steps=WARM_STEPS + (b * PERIOD) // NB, want_obs=False, want_action=False, want_valid=False,
                   envs=(b * E, (b + 1) * E))
nxt = vec.agent_step_rotation("bestfit", E, NB, first_batch=0)
torch.cuda.synchronize()
lib = nv.lib()
buf = (C.c_ulonglong * 16)()
lib.vmgym_debug_prof(buf)                       # read + reset
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nxt = vec.agent_step_rotation("bestfit", E, K, first_batch=nxt)
e1.record()
torch.cuda.synchronize()
lib.vmgym_debug_prof(buf)
names = ["record load wait", "setup", "agent act", "env step (total)", "outputs / obs", "write-back", "  apply", "  arrival draw", "  departures",
         "  clamp + admissions", "  reward + counters", "  act: team scans", "  act: NUMBER of chunk visits", "  act: fit-table builds",
         "  act: candidate filter", "  act: chunk visits (incl. scans)"]
n_rep = E * K if SHAPE == "s1000" else (E + 6) // 7 * K      # reporting warps x items (team mode: every env's main warp; else one warp per CTA of 7)
tot = sum(buf[i] for i in (0, 1, 2, 3, 4, 5))
for i, n in enumerate(names):
    print(f"{n:24s} {buf[i] / n_rep:10.0f} cycles per env-step  {100 * buf[i] / tot:5.1f} %")
print(f"sum of top-level sections {tot / n_rep:10.0f} cycles per env-step = {tot / n_rep / 1.965e3:.2f} us at 1965 MHz; launch: {e0.elapsed_time(e1) * 1e3 / K:.2f} us per batch step")
