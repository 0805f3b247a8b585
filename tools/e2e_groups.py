"""e2e loop (HostVecEnv.run_pipelined) vs number of stream groups, resident / re-uploaded observations, blocking vs polling scheduler.
Every variant is timed twice, alternating, after a long warm-up (the first hundred steps of a fresh process run at lower clocks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np, torch
from bench import WARM_STEPS, load_env_cfg
from vmgym import Config
from vmgym.host_vec import HostVecEnv
cfg = load_env_cfg(); E = 4096
S = 200


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / S * 1e6


for resident, dma in ((True, True), (True, "nofuse")):
    for groups in (2, 4, 8):
        hv = HostVecEnv(Config(**cfg), E, groups=groups, rng="philox", agent="bestfit", resident_obs=resident, action_dma=True, fused_next=(dma is True))
        hv.fast_forward(WARM_STEPS); hv.run_pipelined(300); hv.run_pipelined(300, poll=True)

        def plain():
            for _ in range(S):
                hv.act(); hv.step()
        r = []
        for _ in range(2):
            r.append((timed(lambda: hv.run_pipelined(S)), timed(lambda: hv.run_pipelined(S, poll=True)), timed(plain)))
        best = [min(x[i] for x in r) for i in range(3)]
        print(f"resident={resident} dma={dma} groups={groups}: blocking round-robin {r[0][0]:6.1f} / {r[1][0]:6.1f}  polling {r[0][1]:6.1f} / {r[1][1]:6.1f}  "
              f"plain loop {r[0][2]:6.1f} / {r[1][2]:6.1f} us per step -> best {E / min(best):.2f} M env-steps/s")
        hv.close(); del hv
