"""ncu target: the benchmark's rotation launch (vmgym_agent_step_rotation) in isolation.
    python tools/prof_rotation.py [envs_per_batch] [batches] [K] [replays] [s100|s1000]
Launch order of step_kernel: `batches` warm-up launches, one full untimed rotation, then `replays` launches of K batch steps
(ncu: -k regex:step_kernel -s <batches + 1> -c <replays>)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import PERIOD, WARM_STEPS, load_env_cfg  # noqa: E402
from vmgym import Config, VecVmEnv  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 20
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
R = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = load_env_cfg()
if len(sys.argv) > 5 and sys.argv[5] == "s1000":      # BASELINE config 5's shape: team-mode kernel
    cfg = dict(cfg, pms=1000, vms=3000, sequence="highuniform", arrival_rate=1.6)
seeds = np.concatenate([cfg["seed"] + b * E + np.arange(E, dtype=np.int64) for b in range(NB)])
vec = VecVmEnv(Config(**cfg), NB * E, rng="philox", seeds=seeds)
for b in range(NB):
    vec.agent_step("bestfit", n_steps=WARM_STEPS + (b * PERIOD) // NB, want_obs=False, want_action=False, want_valid=False,
                   envs=(b * E, (b + 1) * E))
nxt = vec.agent_step_rotation("bestfit", E, NB, first_batch=0)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(R)]
for r in range(R):
    ev[r][0].record()
    nxt = vec.agent_step_rotation("bestfit", E, K, first_batch=nxt)
    ev[r][1].record()
torch.cuda.synchronize()
print("rotation launch: %d envs x %d batches, K = %d: us per batch step %s" % (E, NB, K, [round(a.elapsed_time(b) * 1e3 / K, 2) for a, b in ev]))
