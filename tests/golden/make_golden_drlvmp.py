#!/usr/bin/env python
"""Golden vectors for the DRL-VMP rollout side, from the UNMODIFIED reference (src/agents/drlvmp.py):
  * the four placement heuristics `_get_{worstfit,dot,norm2,bestfit}_action` (:549-617) on observations taken from
    reference first-fit runs (the methods only touch `self.env.config`, so they are called on a stand-in `self`);
  * the dueling C51 / NoisyNet `Network` (:326-379): a small random instance's state_dict, inputs and q-values.
Run in the build container only (needs /root/reference):  python tests/golden/make_golden_drlvmp.py
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VMGYM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), REF]
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

from vmenv.envs.env import VmEnv  # noqa: E402
from vmenv.envs.config import Config  # noqa: E402
from src.agents.firstfit import FirstFitAgent  # noqa: E402
import src.agents.drlvmp as ref  # noqa: E402

out = {}
for base, steps, sample_at in (("10", 900, (150, 400, 650, 899)), ("100", 260, (60, 140, 259))):
    cfg = yaml.safe_load(open(os.path.join(REF, "config", f"{base}.yml")))["environment"]
    cfg["reward_function"] = "wr"
    if base == "10":
        cfg.update(arrival_rate=0.25, service_length=60)
    env = VmEnv(Config(**cfg))
    ff = FirstFitAgent(env)
    fake = types.SimpleNamespace(env=env)
    obs, _ = env.reset(seed=cfg["seed"])
    rng = np.random.default_rng(5)
    rows_obs, rows_v, rows_choice, rows_pm = [], [], [], []
    for t in range(steps):
        a = ff.act(obs)
        # leave roughly a third of the proposals unplaced so waiting VMs that fit somewhere exist
        a = np.where(rng.random(a.size) < 0.35, env.vm_placement, a)
        obs, *_ = env.step(a)
        if t in sample_at:
            waiting = np.flatnonzero(env.vm_placement == env.WAIT_STATUS)
            for v in waiting[:10]:
                for choice, fn in enumerate((ref.DRLVMPAgent._get_worstfit_action, ref.DRLVMPAgent._get_dot_action,
                                             ref.DRLVMPAgent._get_norm2_action, ref.DRLVMPAgent._get_bestfit_action)):
                    o = torch.from_numpy(obs.copy()).float()
                    _, action = fn(fake, o, int(v))
                    rows_obs.append(obs.copy()); rows_v.append(int(v)); rows_choice.append(choice)
                    pm = int(action[int(v)])
                    rows_pm.append(pm if pm != env.WAIT_STATUS else -1)
    out[f"h{base}_obs"] = np.array(rows_obs, np.float32)
    out[f"h{base}_v"] = np.array(rows_v, np.int32)
    out[f"h{base}_choice"] = np.array(rows_choice, np.int32)
    out[f"h{base}_pm"] = np.array(rows_pm, np.int32)
    print(base, len(rows_v), "heuristic cases;", int((np.array(rows_pm) < 0).sum()), "without a fitting PM")

torch.manual_seed(3)
support = torch.linspace(0.0, 200.0, 51)
net = ref.Network(110, 24, 4, 51, support)
x = torch.rand(5, 110) * 3
with torch.no_grad():
    q = net(x)
for k, v in net.state_dict().items():
    out["net." + k] = v.numpy()
out["net_x"] = x.numpy()
out["net_q"] = q.numpy()
# ---- prioritized-replay segment trees (src/segment_tree.py) -------------------------------------------------------
from src.segment_tree import MinSegmentTree, SumSegmentTree  # noqa: E402

cap = 1024
st, mt = SumSegmentTree(cap), MinSegmentTree(cap)
rng = np.random.default_rng(11)
b_idx, b_val, sums, mins, ubs, rets = [], [], [], [], [], []
for b in range(12):
    idx = rng.integers(0, 700, size=40)                 # duplicates inside a batch on purpose (last write wins)
    val = rng.random(40) ** 0.6 + 1e-6
    for i, v in zip(idx, val):
        st[int(i)] = float(v)
        mt[int(i)] = float(v)
    total = st.sum()
    ub = rng.random(30) * total
    b_idx.append(idx); b_val.append(val); sums.append(total); mins.append(mt.min()); ubs.append(ub)
    rets.append([st.retrieve(float(u)) for u in ub])
out.update(st_cap=cap, st_idx=np.array(b_idx, np.int64), st_val=np.array(b_val), st_sum=np.array(sums), st_min=np.array(mins),
           st_ub=np.array(ubs), st_ret=np.array(rets, np.int64))
print("segment tree: 12 batches, final sum", sums[-1])

np.savez_compressed(os.path.join(HERE, "drlvmp.npz"), **out)
print("wrote drlvmp.npz", os.path.getsize(os.path.join(HERE, "drlvmp.npz")) // 1024, "KiB")
