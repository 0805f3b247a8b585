"""Fused best-fit / first-fit step time at the three benchmark shapes (one launch per step, observation written)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np, torch, yaml
from bench import WARM_STEPS, load_env_cfg
from vmgym import Config, VecVmEnv
cfg = load_env_cfg()
shapes = {"s1000": (dict(cfg, pms=1000, vms=3000, sequence="highuniform", arrival_rate=1.6), 1024, "bestfit"),
          "s100": (cfg, 4096, "bestfit")}
cfg10 = yaml.safe_load(open(os.path.join(ROOT, "configs", "10.yml")))["environment"]; cfg10["reward_function"] = "wr"
shapes["s10"] = (cfg10, 1 << 20, "firstfit")
only = sys.argv[1:] or list(shapes)
for name in only:
    kw, E, agent = shapes[name]
    v = VecVmEnv(Config(**kw), E, rng="philox")
    v.agent_step(agent, n_steps=WARM_STEPS, want_obs=False, want_action=False, want_valid=False)
    v.agent_step(agent, 1, want_obs=True, want_action=False, want_valid=False)
    torch.cuda.synchronize()
    for reps, n_steps in ((20, 1), (5, 100)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            v.agent_step(agent, n_steps, want_obs=True, want_action=False, want_valid=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name}: {E} envs, {n_steps} step(s)/launch: {ms * 1e3 / n_steps:9.2f} us/step = {E * n_steps / ms / 1e3:9.2f} M env-steps/s")
    del v
