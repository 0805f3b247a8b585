"""GPU-side latency of the pieces of one HostVecEnv group step (CUDA events on the group's stream, ops issued one by one)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), ROOT]
import numpy as np, torch
from bench import WARM_STEPS, load_env_cfg
from vmgym import Config
from vmgym.host_vec import HostVecEnv
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = load_env_cfg(); E = 4096
hv = HostVecEnv(Config(**cfg), E, groups=G, rng="philox", agent="bestfit", use_graphs=False)
hv.fast_forward(WARM_STEPS + 500); hv.run_pipelined(50)
g = hv.groups[0]
n = g.hi - g.lo


def t(name, fn, reps=50):
    with torch.cuda.stream(g.stream):
        fn(); g.stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(reps):
            e0.record(g.stream); fn(); e1.record(g.stream); g.stream.synchronize()
            tot += e0.elapsed_time(e1)
    print(f"{name:44s} {tot / reps * 1e3:8.1f} us")


t("H2D actions (%d x %d B)" % (n, hv.V), lambda: g.d_act_in.copy_(hv.action[g.lo:g.hi], non_blocking=True))
t("D2H actions", lambda: hv.action[g.lo:g.hi].copy_(g.d_act_in, non_blocking=True))
t("agent.act on device obs -> device actions", lambda: g.agent.act(g.vec.obs, out=g.d_act_in))
t("vec.step (device actions, host mirror outputs)", lambda: g.vec.step(g.d_act_in, want_valid=False, obs_mirror=hv.obs[g.lo:g.hi],
                                                                     host_outputs=(hv.reward[g.lo:g.hi], hv.terminated[g.lo:g.hi])))
t("vec.step (device actions, device outputs)", lambda: g.vec.step(g.d_act_in, want_valid=False))
t("empty (event pair)", lambda: None)
