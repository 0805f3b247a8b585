"""DRL-VMP agent, rollout side (src/agents/drlvmp.py:326-379 network, :504-617 act + heuristics): a dueling C51 network
with NoisyNet heads chooses, for every waiting VM in slot order, one of four placement heuristics; the chosen PM is
written into the observation copy (PM loads are NOT updated, drlvmp.py:557-565) and the loop continues on it.

State-dict keys mirror the reference's modules (`feature_layer.0.*`, `advantage_hidden_layer.weight_mu` ...), with or
without the `_orig_mod.` prefix of torch.compile, so reference checkpoints load.  The heuristic selection runs in the
`vmgym_drlvmp_choice` kernel; the network layers run through torch.  Training internals (n-step buffer, prioritized
replay on the device segment trees, C51 projection kernel, batched learn loop) live in vmgym/drlvmp_train.py."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as nv


@dataclass
class DRLVMPConfig:
    episodes: int = 2000
    hidden_size: int = 256
    lr: float = 3e-5
    memory_size: int = 100000
    batch_size: int = 100
    target_update: int = 5
    gamma: float = 0.99
    alpha: float = 0.2
    beta: float = 0.5
    prior_eps: float = 1e-6
    v_min: float = 0.0
    v_max: float = 200.0
    atom_size: int = 51
    n_step: int = 3
    device: str = "cuda"
    show_training_progress: bool = True


class NoisyLinear(nn.Module):
    """drlvmp.py:243-324: y = x (W_mu + W_sigma * W_eps)^T + (b_mu + b_sigma * b_eps), factorised Gaussian noise."""

    def __init__(self, in_features: int, out_features: int, std_init: float = 0.5):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        r = 1.0 / math.sqrt(in_features)
        self.weight_mu = nn.Parameter(torch.empty(out_features, in_features).uniform_(-r, r))
        self.weight_sigma = nn.Parameter(torch.full((out_features, in_features), std_init / math.sqrt(in_features)))
        self.register_buffer("weight_epsilon", torch.zeros(out_features, in_features))
        self.bias_mu = nn.Parameter(torch.empty(out_features).uniform_(-r, r))
        self.bias_sigma = nn.Parameter(torch.full((out_features,), std_init / math.sqrt(out_features)))
        self.register_buffer("bias_epsilon", torch.zeros(out_features))
        self.reset_noise()

    @staticmethod
    def _scaled(n):
        x = torch.randn(n)
        return x.sign() * x.abs().sqrt()

    def reset_noise(self):
        e_in, e_out = self._scaled(self.in_features), self._scaled(self.out_features)
        self.weight_epsilon.copy_(torch.outer(e_out, e_in))
        self.bias_epsilon.copy_(e_out)

    def forward(self, x):
        return F.linear(x, self.weight_mu + self.weight_sigma * self.weight_epsilon,
                        self.bias_mu + self.bias_sigma * self.bias_epsilon)


class Network(nn.Module):
    """drlvmp.py:326-379: Linear+ReLU feature -> noisy advantage / value heads over `atom_size` atoms -> dueling
    combination -> softmax (clamped at 1e-3) -> expected value over the support."""

    def __init__(self, in_dim: int, hidden_size: int, out_dim: int, atom_size: int, support: torch.Tensor):
        super().__init__()
        self.out_dim, self.atom_size = out_dim, atom_size
        self.register_buffer("support", support.clone(), persistent=False)
        self.feature_layer = nn.Sequential(nn.Linear(in_dim, hidden_size), nn.ReLU())
        self.advantage_hidden_layer = NoisyLinear(hidden_size, hidden_size)
        self.advantage_layer = NoisyLinear(hidden_size, out_dim * atom_size)
        self.value_hidden_layer = NoisyLinear(hidden_size, hidden_size)
        self.value_layer = NoisyLinear(hidden_size, atom_size)

    def dist(self, x):
        f = self.feature_layer(x)
        adv = self.advantage_layer(F.relu(self.advantage_hidden_layer(f))).view(-1, self.out_dim, self.atom_size)
        val = self.value_layer(F.relu(self.value_hidden_layer(f))).view(-1, 1, self.atom_size)
        q_atoms = val + adv - adv.mean(dim=1, keepdim=True)
        return F.softmax(q_atoms, dim=-1).clamp(min=1e-3)

    def forward(self, x):
        return torch.sum(self.dist(x) * self.support, dim=2)

    def reset_noise(self):
        for m in (self.advantage_hidden_layer, self.advantage_layer, self.value_hidden_layer, self.value_layer):
            m.reset_noise()


class DRLVMPAgent:
    name = "DRLVMPAgent"

    def __init__(self, env, config: DRLVMPConfig | None = None):
        self.env = env
        self.vec = getattr(env, "vec", env)
        self.config = config or DRLVMPConfig()
        vec = self.vec
        self.device = vec.device
        self.n_actions = 4
        self.support = torch.linspace(self.config.v_min, self.config.v_max, self.config.atom_size, device=self.device)
        self.dqn = Network(vec.obs_dim, self.config.hidden_size, self.n_actions, self.config.atom_size, self.support).to(self.device)

    def eval(self, mode=True):
        self.dqn.train(not mode)

    def learn(self, episodes: int | None = None, max_steps: int | None = None, updates_per_step: int = 1):
        """drlvmp.py:433-500 over all envs of the batch (vmgym/drlvmp_train.py)."""
        from .drlvmp_train import DRLVMPTrainer
        self.trainer = DRLVMPTrainer(self, updates_per_step=updates_per_step)
        return self.trainer.learn(episodes=episodes, max_steps=max_steps)

    def save_model(self, modelpath):
        if modelpath:
            torch.save({"_orig_mod." + k: v for k, v in self.dqn.state_dict().items()}, modelpath)

    def load_model(self, modelpath):
        sd = torch.load(modelpath, map_location=self.device)
        self.dqn.load_state_dict({k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k: v for k, v in sd.items()})
        self.dqn.eval()

    def heuristic(self, obs, vm_index, choice):
        """_convert_action (drlvmp.py:517-530) for one VM per env: returns the chosen PM [n] (int32, -1 = none fits)."""
        vec = self.vec
        n = obs.shape[0]
        pm = torch.empty(n, dtype=torch.int32, device=self.device)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(nv.lib().vmgym_drlvmp_choice(C.byref(vec._ccfg()), obs.data_ptr(), vm_index.data_ptr(), choice.data_ptr(), n,
                                              pm.data_ptr(), stream), "vmgym_drlvmp_choice")
        return pm

    @torch.no_grad()
    def act(self, observation):
        """drlvmp.py:504-512 for a batch of observations [n, D] (or one numpy observation)."""
        vec = self.vec
        host = isinstance(observation, np.ndarray)
        obs = torch.from_numpy(np.ascontiguousarray(observation, np.float32)).to(self.device) if host else observation
        single = obs.dim() == 1
        obs = obs.reshape(-1, vec.obs_dim).float().contiguous().clone()       # the working observation copy
        n, V, P = obs.shape[0], vec.V, vec.P
        waiting = obs[:, :V] == float(P)                                       # fixed at entry (drlvmp.py:507)
        n_wait = waiting.sum(1)
        # k-th waiting slot of every env, in slot order
        order = torch.argsort((~waiting).to(torch.int8), dim=1, stable=True).to(torch.int32)
        rows = torch.arange(n, device=self.device)
        for k in range(int(n_wait.max().item()) if n else 0):
            active = n_wait > k
            idx = torch.where(active, order[:, k], torch.full_like(order[:, k], -1)).contiguous()
            choice = self.dqn(obs).argmax(dim=1).to(torch.int32).contiguous()   # drlvmp.py:514-515
            pm = self.heuristic(obs, idx, choice)
            upd = active & (pm >= 0)
            obs[rows[upd], idx[upd].long()] = pm[upd].float()                   # placement written into the obs copy only
        action = obs[:, :V].to(torch.int64)
        if host:
            a = action.cpu().numpy()
            return a[0] if single else a
        return action[0] if single else action


class DeviceSegmentTrees:
    """Sum + Min segment trees of the prioritized replay buffer (src/segment_tree.py, drlvmp.py:157-241) on the device:
    batched `tree[idx] = priority ** alpha` stores, `sum()`, `min()` and batched `retrieve(upperbound)` with the
    reference's fp64 association (bit-identical to its Python floats)."""

    def __init__(self, capacity: int, device="cuda"):
        assert capacity > 0 and capacity & (capacity - 1) == 0, "capacity must be positive and a power of 2."   # :30-32
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.sum_tree = torch.zeros(2 * capacity, dtype=torch.float64, device=self.device)                     # init 0.0
        self.min_tree = torch.full((2 * capacity,), float("inf"), dtype=torch.float64, device=self.device)     # init inf

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set(self, idx, val):
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).contiguous()
        val = torch.as_tensor(val, dtype=torch.float64, device=self.device).contiguous()
        nv.check(nv.lib().vmgym_segtree_update(self.sum_tree.data_ptr(), self.min_tree.data_ptr(), self.capacity, idx.data_ptr(),
                                               val.data_ptr(), idx.numel(), self._stream()), "vmgym_segtree_update")

    def sum(self):
        return self.sum_tree[1]

    def min(self):
        return self.min_tree[1]

    def retrieve(self, upperbound):
        ub = torch.as_tensor(upperbound, dtype=torch.float64, device=self.device).contiguous()
        out = torch.empty(ub.numel(), dtype=torch.int64, device=self.device)
        nv.check(nv.lib().vmgym_segtree_retrieve(self.sum_tree.data_ptr(), self.capacity, ub.data_ptr(), ub.numel(),
                                                 out.data_ptr(), self._stream()), "vmgym_segtree_retrieve")
        return out
