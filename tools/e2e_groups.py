"""e2e loop (HostVecEnv.run_pipelined) vs number of stream groups, resident / re-uploaded observations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np, torch
from bench import WARM_STEPS, load_env_cfg
from vmgym import Config
from vmgym.host_vec import HostVecEnv
cfg = load_env_cfg(); E = 4096
for resident, dma in ((True, True), (True, False), (False, True)):
    for groups in (1, 2, 4, 8, 16):
        hv = HostVecEnv(Config(**cfg), E, groups=groups, rng="philox", agent="bestfit", resident_obs=resident, action_dma=dma)
        hv.fast_forward(WARM_STEPS); hv.run_pipelined(5); torch.cuda.synchronize()
        t0 = time.perf_counter(); hv.run_pipelined(100); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        for _ in range(50):
            hv.act(); hv.step()
        torch.cuda.synchronize(); dt2 = time.perf_counter() - t1
        print(f"resident={resident} action_dma={dma} groups={groups}: pipelined {dt / 100 * 1e6:7.1f} us/step = {E * 100 / dt / 1e6:6.2f} M env-steps/s; plain loop {dt2 / 50 * 1e6:7.1f} us/step")
        hv.close(); del hv
