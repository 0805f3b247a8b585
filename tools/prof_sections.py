"""Cycles per section of the step kernel's decision warp (needs a library built with VMGYM_NVCC_EXTRA=-DVMGYM_PROF).
    python tools/prof_sections.py s1000|s100 [steps]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np, torch
from bench import WARM_STEPS, load_env_cfg
from vmgym import Config, VecVmEnv
from vmgym import _native as nv
cfg = load_env_cfg()
name = sys.argv[1] if len(sys.argv) > 1 else "s1000"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
agent = "bestfit"
if name == "s1000":
    kw, E = dict(cfg, pms=1000, vms=3000, sequence="highuniform", arrival_rate=1.6), 1024
elif name == "s10":
    import yaml
    kw = yaml.safe_load(open(os.path.join(ROOT, "configs", "10.yml")))["environment"]
    kw["reward_function"] = "wr"
    E, agent = 1 << 20, "firstfit"
else:
    kw, E = cfg, 4096
v = VecVmEnv(Config(**kw), E, rng="philox")
v.agent_step(agent, n_steps=WARM_STEPS, want_obs=False, want_action=False, want_valid=False)
lib = nv.lib()
buf = (C.c_ulonglong * 16)()
lib.vmgym_debug_prof(buf)
for _ in range(steps):
    v.agent_step(agent, 1, want_obs=True, want_action=False, want_valid=False)
lib.vmgym_debug_prof(buf)
names = ["record load wait", "setup", "agent act", "env step (total)", "outputs / obs", "write-back", "  apply", "  arrival draw", "  departures",
         "  clamp + admissions", "  reward + counters", "  act: team scans", "  act: NUMBER of chunk visits", "  act: fit-table builds", "  act: candidate filter",
         "  act: chunk visits (incl. scans)"]
tot = sum(buf[i] for i in (0, 1, 2, 3, 4, 5))
n_rep = (E * steps) if name == "s1000" else (E * steps) / int(os.environ.get("PROF_WARPS_PER_CTA", "4"))   # thread 0 of each CTA reports
for i, n in enumerate(names):
    print(f"{n:24s} {buf[i] / n_rep:10.0f} cycles per env-step  {100 * buf[i] / tot:5.1f} %")
print(f"sum of top-level sections {tot / n_rep:10.0f} cycles per reported env-step = {tot / n_rep / 1.965e3:.2f} us at 1965 MHz (warps per CTA assumed {os.environ.get('PROF_WARPS_PER_CTA', '4')})")
