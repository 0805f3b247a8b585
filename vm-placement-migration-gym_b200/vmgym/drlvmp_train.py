"""DRL-VMP training internals on the device (src/agents/drlvmp.py:21-241, 450-500, 618-710): n-step transition builder,
prioritized replay on the device segment trees, the categorical-DQN (C51) loss with the projection kernel, and the
training loop over a BATCH of envs.

The reference trains on one env: one transition and one optimisation step per env step.  Here N envs step together;
every env step contributes N transitions (stored in env order, exactly what N interleaved reference streams would
store) and is followed by `updates_per_step` optimisation steps on `batch_size` prioritized samples.  With N = 1 and
the reference's uniforms the buffers evolve exactly like the reference's (tests/test_drlvmp.py, golden vectors from
the reference classes); training results themselves are parity-unpinned (the reference ships no DRL-VMP weights)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nv
from .drlvmp import DeviceSegmentTrees, Network


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class NStepReplay:
    """ReplayBuffer with n-step returns (drlvmp.py:21-116) for N parallel envs.  Each env has its own sliding window of
    the last n transitions (the reference's deque, which — like the reference — is NOT cleared at episode ends); when a
    window is full, the n-step transition (:103-113) of every env is appended to the shared ring in env order."""

    def __init__(self, obs_dim: int, size: int, num_envs: int = 1, n_step: int = 1, gamma: float = 0.99, device="cuda"):
        self.device = torch.device(device)
        self.obs_buf = torch.zeros((size, obs_dim), dtype=torch.float32, device=self.device)
        self.next_obs_buf = torch.zeros((size, obs_dim), dtype=torch.float32, device=self.device)
        self.acts_buf = torch.zeros(size, dtype=torch.int32, device=self.device)
        self.rews_buf = torch.zeros(size, dtype=torch.float32, device=self.device)
        self.done_buf = torch.zeros(size, dtype=torch.int32, device=self.device)
        self.max_size, self.ptr, self.size = int(size), 0, 0
        self.n_step, self.gamma, self.N = int(n_step), float(gamma), int(num_envs)
        n, N = self.n_step, self.N
        self.w_obs = torch.zeros((n, N, obs_dim), dtype=torch.float32, device=self.device)
        self.w_next = torch.zeros((n, N, obs_dim), dtype=torch.float32, device=self.device)
        self.w_act = torch.zeros((n, N), dtype=torch.int32, device=self.device)
        self.w_rew = torch.zeros((n, N), dtype=torch.float64, device=self.device)
        self.w_done = torch.zeros((n, N), dtype=torch.bool, device=self.device)
        self.filled = 0                   # transitions seen per env (the same for all envs)

    def __len__(self):
        return self.size

    def store(self, obs, act, rew, next_obs, done):
        """One transition per env ([N, D], [N], [N], [N, D], [N]).  Returns None until the windows are full, else the
        ring positions [N] the n-step transitions were written to and the windows' oldest 1-step transitions
        (obs, act, rew, next_obs, done) — what ReplayBuffer.store returns (:74)."""
        n = self.n_step
        slot = self.filled % n
        self.w_obs[slot].copy_(obs); self.w_next[slot].copy_(next_obs)
        self.w_act[slot].copy_(act.to(torch.int32)); self.w_rew[slot].copy_(rew.to(torch.float64))
        self.w_done[slot].copy_(done.to(torch.bool))
        self.filled += 1
        if self.filled < n:
            return None
        order = [(self.filled - n + k) % n for k in range(n)]             # oldest ... newest
        last = order[-1]
        rew_n = self.w_rew[last].clone()
        next_n = self.w_next[last].clone()
        done_n = self.w_done[last].clone()
        for k in reversed(order[:-1]):                                     # :107-111
            d = self.w_done[k]
            rew_n = self.w_rew[k] + self.gamma * rew_n * (1.0 - d.to(torch.float64))
            next_n = torch.where(d.unsqueeze(1), self.w_next[k], next_n)
            done_n = torch.where(d, d, done_n)
        first = order[0]
        pos = (self.ptr + torch.arange(self.N, device=self.device)) % self.max_size
        self.obs_buf[pos] = self.w_obs[first]
        self.next_obs_buf[pos] = next_n
        self.acts_buf[pos] = self.w_act[first]
        self.rews_buf[pos] = rew_n.to(torch.float32)
        self.done_buf[pos] = done_n.to(torch.int32)
        self.ptr = (self.ptr + self.N) % self.max_size
        self.size = min(self.size + self.N, self.max_size)
        return pos, (self.w_obs[first], self.w_act[first], self.w_rew[first], self.w_next[first], self.w_done[first])

    def sample_batch_from_idxs(self, idxs):
        return dict(obs=self.obs_buf[idxs], next_obs=self.next_obs_buf[idxs], acts=self.acts_buf[idxs], rews=self.rews_buf[idxs],
                    done=self.done_buf[idxs])


class PrioritizedReplay(NStepReplay):
    """PrioritizedReplayBuffer (drlvmp.py:118-241) with the sum / min trees on the device."""

    def __init__(self, obs_dim: int, size: int, num_envs: int = 1, alpha: float = 0.6, n_step: int = 1, gamma: float = 0.99,
                 device="cuda"):
        assert alpha >= 0
        super().__init__(obs_dim, size, num_envs, n_step, gamma, device)
        self.alpha = float(alpha)
        self.max_priority = torch.ones((), dtype=torch.float64, device=self.device)
        self.tree_ptr = 0
        cap = 1
        while cap < self.max_size:
            cap *= 2
        self.trees = DeviceSegmentTrees(cap, self.device)

    def store(self, obs, act, rew, next_obs, done):
        out = super().store(obs, act, rew, next_obs, done)
        if out is not None:                                                # :170-173
            idx = (self.tree_ptr + torch.arange(self.N, device=self.device)) % self.max_size
            self.trees.set(idx, (self.max_priority ** self.alpha).expand(self.N))
            self.tree_ptr = (self.tree_ptr + self.N) % self.max_size
        return out

    def sample_batch(self, batch_size: int, beta: float = 0.4, u=None):
        """:178-203.  `u`: optional float64 uniforms [batch_size] (default: torch.rand on the device)."""
        assert len(self) >= batch_size and beta > 0
        if u is None:
            u = torch.rand(batch_size, dtype=torch.float64, device=self.device)
        u = torch.as_tensor(u, dtype=torch.float64, device=self.device).contiguous()
        idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        w = torch.empty(batch_size, dtype=torch.float64, device=self.device)
        t = self.trees
        nv.check(nv.lib().vmgym_per_sample(t.sum_tree.data_ptr(), t.min_tree.data_ptr(), t.capacity, len(self), batch_size,
                                           u.data_ptr(), float(beta), idx.data_ptr(), w.data_ptr(), _stream(self.device)),
                 "vmgym_per_sample")
        d = self.sample_batch_from_idxs(idx)
        d.update(weights=w, indices=idx)
        return d

    def update_priorities(self, indices, priorities):
        """:205-215 (duplicated indices: the last priority wins, like the sequential loop)."""
        priorities = torch.as_tensor(priorities, dtype=torch.float64, device=self.device)
        self.trees.set(indices, priorities ** self.alpha)
        self.max_priority = torch.maximum(self.max_priority, priorities.max())


def c51_project(next_dist, reward, done, support, gamma: float, v_min: float, v_max: float):
    """The projected target distribution of drlvmp.py:676-699 (kernel `vmgym_c51_project`)."""
    n, atoms = next_dist.shape
    next_dist = next_dist.contiguous().float()
    reward = reward.reshape(-1).contiguous().float()
    done = done.reshape(-1).to(torch.int32).contiguous()
    support = support.contiguous().float()
    proj = torch.empty_like(next_dist)
    nv.check(nv.lib().vmgym_c51_project(next_dist.data_ptr(), reward.data_ptr(), done.data_ptr(), support.data_ptr(), float(gamma),
                                        float(v_min), float(v_max), atoms, n, proj.data_ptr(), _stream(next_dist.device)),
             "vmgym_c51_project")
    return proj


def dqn_loss(dqn: Network, dqn_target: Network, samples, gamma: float, v_min: float, v_max: float):
    """_compute_dqn_loss (drlvmp.py:661-706): double-DQN action from the online net, its distribution under the target
    net, C51 projection, cross entropy against the online distribution of the taken action.  Returns [B]."""
    state, next_state = samples["obs"], samples["next_obs"]
    action = samples["acts"].long()
    B = state.shape[0]
    rows = torch.arange(B, device=state.device)
    with torch.no_grad():
        next_action = dqn(next_state).argmax(1)
        next_dist = dqn_target.dist(next_state)[rows, next_action]
        proj = c51_project(next_dist, samples["rews"], samples["done"], dqn.support, gamma, v_min, v_max)
    log_p = torch.log(dqn.dist(state)[rows, action])
    return -(proj * log_p).sum(1)


class DRLVMPTrainer:
    """DRLVMPAgent.learn (drlvmp.py:433-500) over the envs of a VecVmEnv."""

    def __init__(self, agent, updates_per_step: int = 1):
        self.agent = agent
        cfg, vec = agent.config, agent.vec
        self.vec, self.cfg, self.device = vec, cfg, agent.device
        N, D = vec.num_envs, vec.obs_dim
        self.memory = PrioritizedReplay(D, cfg.memory_size, N, alpha=cfg.alpha, device=self.device)
        self.use_n_step = cfg.n_step > 1
        self.memory_n = NStepReplay(D, cfg.memory_size, N, n_step=cfg.n_step, gamma=cfg.gamma, device=self.device) if self.use_n_step else None
        self.dqn = agent.dqn
        self.dqn_target = Network(D, cfg.hidden_size, agent.n_actions, cfg.atom_size, agent.support).to(self.device)
        self.dqn_target.load_state_dict(self.dqn.state_dict())
        self.dqn_target.eval()
        self.optimizer = torch.optim.Adam(self.dqn.parameters(), lr=cfg.lr)
        self.beta = float(cfg.beta)
        self.updates_per_step = int(updates_per_step)
        self.update_cnt = 0
        self.losses = []

    def optimize(self):
        """_optimize_model (:618-659)."""
        cfg = self.cfg
        samples = self.memory.sample_batch(cfg.batch_size, self.beta)
        weights = samples["weights"].float().reshape(-1, 1)
        idx = samples["indices"]
        elementwise = dqn_loss(self.dqn, self.dqn_target, samples, cfg.gamma, cfg.v_min, cfg.v_max)
        if self.use_n_step:
            elementwise = elementwise + dqn_loss(self.dqn, self.dqn_target, self.memory_n.sample_batch_from_idxs(idx),
                                                 cfg.gamma ** cfg.n_step, cfg.v_min, cfg.v_max)
        loss = torch.mean(elementwise * weights)          # broadcasts [B] x [B, 1] -> [B, B] exactly like the reference (:628,640)
        self.optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.dqn.parameters(), 10.0)
        self.optimizer.step()
        self.memory.update_priorities(idx, elementwise.detach().double() + cfg.prior_eps)
        self.dqn.reset_noise()
        self.dqn_target.reset_noise()
        return loss

    def env_action(self, obs):
        """One decision per env and step, as in learn() (:455-462): the online net picks a heuristic for the FIRST waiting
        VM; every other slot keeps its placement.  Returns (choice [N] int32, env action [N, V] placement dtype)."""
        vec, agent = self.vec, self.agent
        V, P = vec.V, vec.P
        with torch.no_grad():
            choice = self.dqn(obs).argmax(dim=1).to(torch.int32).contiguous()
        placement = obs[:, :V]
        waiting = placement == float(P)
        has = waiting.any(dim=1)
        first = torch.where(has, waiting.to(torch.int8).argmax(dim=1), torch.full((obs.shape[0],), -1, device=obs.device)).to(torch.int32)
        pm = agent.heuristic(obs.contiguous(), first.contiguous(), choice)
        action = placement.to(torch.int64)
        rows = torch.nonzero(has & (pm >= 0)).flatten()
        action[rows, first[rows].long()] = pm[rows].long()
        return choice, action

    def learn(self, episodes: int | None = None, max_steps: int | None = None):
        cfg, vec = self.cfg, self.vec
        episodes = int(cfg.episodes if episodes is None else episodes)
        returns = np.zeros((episodes, vec.num_envs))
        steps = 0
        self.dqn.train()
        for ep in range(episodes):
            vec.seed(vec.config.seed + ep * vec.num_envs + np.arange(vec.num_envs))      # :450 a different sequence per episode
            obs, _ = vec.reset()
            obs = obs.clone()
            ep_ret = torch.zeros(vec.num_envs, dtype=torch.float64, device=self.device)
            done = False
            while not done:
                choice, action = self.env_action(obs)
                nobs, reward, term, trunc, _ = vec.step(action, want_valid=False)
                nobs = nobs.clone()
                fraction = min(ep / max(1, cfg.episodes), 1.0)
                self.beta = self.beta + fraction * (1.0 - self.beta)                      # :467-468
                tr = (obs, choice, reward, nobs, term)
                if self.use_n_step:
                    out = self.memory_n.store(*tr)
                    one_step = out[1] if out is not None else None
                else:
                    one_step = tr
                if one_step is not None:
                    self.memory.store(*one_step)
                if len(self.memory) >= cfg.batch_size:
                    for _ in range(self.updates_per_step):
                        self.losses.append(self.optimize())
                        self.update_cnt += 1
                        if self.update_cnt % cfg.target_update == 0:
                            self.dqn_target.load_state_dict(self.dqn.state_dict())        # :708-710
                obs = nobs
                ep_ret += reward
                steps += 1
                done = bool(term.all().item()) or (max_steps is not None and steps >= max_steps)
            returns[ep] = ep_ret.cpu().numpy()
            if max_steps is not None and steps >= max_steps:
                break
        return returns
