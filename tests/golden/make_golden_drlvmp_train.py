#!/usr/bin/env python
"""Golden vectors for the DRL-VMP TRAINING internals, from the UNMODIFIED reference (src/agents/drlvmp.py):
  * `DRLVMPAgent._compute_dqn_loss` (:661-706: double-DQN action choice, C51 projection with the index_add_ pair, cross
    entropy) on a small agent with distinct online / target weights, 1-step and n-step gammas;
  * `ReplayBuffer.store` / `_get_n_step_info` (:46-113): what lands in the buffers for a stream of transitions with dones;
  * `PrioritizedReplayBuffer` (:118-241): stores, priority updates, `_sample_proportional` (the python `random` draws are
    recorded as uniforms u so that upperbound = a + (b - a) * u can be replayed) and `_calculate_weight`.
torch.compile is switched off through TORCH_COMPILE_DISABLE (an environment switch, the reference code is untouched).
Run in the build container only:  python tests/golden/make_golden_drlvmp_train.py  ->  tests/golden/drlvmp_train.npz
"""
import os
import random
import sys

os.environ["TORCH_COMPILE_DISABLE"] = "1"
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
import src.agents.drlvmp as ref  # noqa: E402

out = {}
cfg = yaml.safe_load(open(os.path.join(REF, "config", "10.yml")))["environment"]
cfg["reward_function"] = "wr"
env = VmEnv(Config(**cfg))
B, H = 12, 24
torch.manual_seed(11)
agent = ref.DRLVMPAgent(env, ref.DRLVMPConfig(hidden_size=H, batch_size=B, memory_size=64, n_step=3, device="cpu",
                                              show_training_progress=False))
# distinct target weights so that the double-DQN argmax (online) and the evaluated distribution (target) differ
tgt = ref.Network(110, H, 4, 51, agent.support)
agent.dqn_target.load_state_dict({"_orig_mod." + k: v for k, v in tgt.state_dict().items()})
D = 110
g = torch.Generator().manual_seed(5)
samples = dict(obs=torch.rand(B, D, generator=g) * 3, next_obs=torch.rand(B, D, generator=g) * 3,
               acts=torch.randint(0, 4, (B,), generator=g).int(),
               # rewards chosen so that some projected atoms land exactly on the grid (b integral: both index_add weights 0)
               rews=torch.tensor([0.0, 4.0, -3.0, 250.0, 1.5, 7.25, 100.0, 0.37, 12.0, 199.0, 3.99, 8.0]),
               done=torch.tensor([0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1]).int())
for k, v in samples.items():
    out["c51_" + k] = v.numpy()
for name, net in (("dqn", agent.dqn), ("tgt", agent.dqn_target)):
    for k, v in net.state_dict().items():
        out[f"c51_{name}." + k.replace("_orig_mod.", "")] = v.detach().numpy().copy()
for tag, gamma in (("g1", 0.99), ("g3", 0.99 ** 3)):
    loss = agent._compute_dqn_loss(samples, gamma)
    out["c51_loss_" + tag] = loss.detach().numpy()
    out["c51_gamma_" + tag] = gamma
print("c51 losses", out["c51_loss_g1"][:4], out["c51_loss_g3"][:4])

# ---- n-step buffer ---------------------------------------------------------------------------------------------
rb = ref.ReplayBuffer(4, 32, batch_size=4, n_step=3, gamma=0.99)
rng = np.random.default_rng(2)
T = 40
obs = rng.random((T + 1, 4)).astype(np.float32)
acts = rng.integers(0, 4, T)
rews = rng.normal(size=T).astype(np.float64)
dones = rng.random(T) < 0.2
for t in range(T):
    rb.store(torch.from_numpy(obs[t]), torch.tensor(int(acts[t])), float(rews[t]), torch.from_numpy(obs[t + 1]), bool(dones[t]))
out.update(ns_obs=obs, ns_acts=acts, ns_rews=rews, ns_dones=dones.astype(np.uint8), ns_ptr=rb.ptr, ns_size=rb.size,
           ns_obs_buf=rb.obs_buf.numpy(), ns_next_obs_buf=rb.next_obs_buf.numpy(), ns_acts_buf=rb.acts_buf.numpy(),
           ns_rews_buf=rb.rews_buf.numpy(), ns_done_buf=rb.done_buf.numpy())
print("n-step: stored", rb.size, "ptr", rb.ptr)

# ---- prioritized replay -----------------------------------------------------------------------------------------
per = ref.PrioritizedReplayBuffer(4, 48, batch_size=8, alpha=0.2)
for t in range(30):
    per.store(torch.from_numpy(obs[t % T]), torch.tensor(1), 0.0, torch.from_numpy(obs[t % T + 1]), False)
upd_idx, upd_pri, samp_u, samp_idx, samp_w = [], [], [], [], []
for r in range(6):
    random.seed(100 + r)
    us = [random.random() for _ in range(8)]        # random.uniform(a, b) = a + (b - a) * random.random()
    random.seed(100 + r)
    s = per.sample_batch(beta=0.5 + 0.05 * r)
    samp_u.append(us); samp_idx.append(s["indices"]); samp_w.append(s["weights"])
    pri = rng.random(8) * 3 + 1e-6
    per.update_priorities(s["indices"], pri)
    upd_idx.append(s["indices"]); upd_pri.append(pri)
out.update(per_n=30, per_alpha=0.2, per_u=np.array(samp_u), per_idx=np.array(samp_idx, np.int64), per_w=np.array(samp_w),
           per_upd_pri=np.array(upd_pri), per_max_priority=per.max_priority, per_sum=per.sum_tree.sum(), per_min=per.min_tree.min())
print("per: final sum", per.sum_tree.sum(), "max priority", per.max_priority)
path = os.path.join(HERE, "drlvmp_train.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
