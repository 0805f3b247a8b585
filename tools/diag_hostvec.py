"""Where does the HostVecEnv loop spend its time?  Chain latencies in isolation, host cost of an enqueue, timeline of a pipelined run."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vm-placement-migration-gym_b200"))
import numpy as np, torch, yaml
from vmgym import Config
from vmgym.host_vec import HostVecEnv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]; cfg["reward_function"] = "wr"
E, G = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4
hv = HostVecEnv(Config(**cfg), E, groups=G, agent="bestfit")
hv.fast_forward(3000); hv.run_pipelined(3); torch.cuda.synchronize()
pc = time.perf_counter
for which in ("act", "step"):
    lat, host = [], []
    for _ in range(20):
        t0 = pc(); (hv.act_async if which == "act" else hv.step_async)(0); t1 = pc()
        (hv.act_wait if which == "act" else hv.step_wait)(0); t2 = pc()
        host.append(t1 - t0); lat.append(t2 - t0)
    print(f"group-0 {which} chain alone ({E // G} envs): enqueue {np.median(host) * 1e6:.0f} us on the host, complete after {np.median(lat) * 1e6:.0f} us")
# timeline of the pipelined loop
log = []
orig_run = hv._run
def traced(g, which):
    t0 = pc(); orig_run(g, which); log.append((t0, pc(), g.lo, which))
hv._run = traced
t0 = pc(); hv.run_pipelined(10); torch.cuda.synchronize(); t1 = pc()
print(f"pipelined: {(t1 - t0) / 10 * 1e3:.3f} ms/step; enqueue calls {len(log)}, host time in enqueues {sum(b - a for a, b, _, _ in log) / 10 * 1e3:.3f} ms/step")
for a, b, lo, which in log[8:8 + 4 * G]:
    print(f"  t={1e6 * (a - t0):8.0f} us  +{1e6 * (b - a):4.0f}  group@{lo:<5d} {which}")
