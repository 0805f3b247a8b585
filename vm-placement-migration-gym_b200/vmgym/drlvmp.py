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
from .agents import AgentBase


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

    def effective_heads(self):
        """(weight, bias) of the four noisy layers with the current noise folded in (constant between reset_noise calls)."""
        return [(m.weight_mu + m.weight_sigma * m.weight_epsilon, m.bias_mu + m.bias_sigma * m.bias_epsilon)
                for m in (self.advantage_hidden_layer, self.advantage_layer, self.value_hidden_layer, self.value_layer)]

    def q_from_pre(self, pre, eff=None):
        """q-values from the feature layer's PRE-activation (x W1^T + b1): lets a caller that changes one input column
        at a time update `pre` by a rank-1 term instead of redoing the D x H product (DRLVMPAgent.act).  `eff`: the
        result of effective_heads(), to fold the noise once per act() call instead of once per waiting VM."""
        f = F.relu(pre)
        if eff is None:
            eff = self.effective_heads()
        (wah, bah), (wa, ba), (wvh, bvh), (wv, bv) = eff
        adv = F.linear(F.relu(F.linear(f, wah, bah)), wa, ba).view(-1, self.out_dim, self.atom_size)
        val = F.linear(F.relu(F.linear(f, wvh, bvh)), wv, bv).view(-1, 1, self.atom_size)
        dist = F.softmax(val + adv - adv.mean(dim=1, keepdim=True), dim=-1).clamp(min=1e-3)
        return torch.sum(dist * self.support, dim=2)

    def reset_noise(self):
        for m in (self.advantage_hidden_layer, self.advantage_layer, self.value_hidden_layer, self.value_layer):
            m.reset_noise()


class DRLVMPAgent(AgentBase):
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
    def act(self, observation, incremental: bool = True, refresh: int = 128, graph: bool | None = None, fused: bool | None = None):
        """drlvmp.py:504-512 for a batch of observations [n, D] (or one numpy observation).

        The reference re-evaluates the whole network for every waiting VM although only ONE observation entry changed
        (the VM's placement, WAIT -> chosen PM).  With `incremental` the feature layer's pre-activation is carried along
        and corrected by that entry's weight column (rank-1 update per env), and recomputed from scratch every `refresh`
        VMs; the difference to the full product is fp32 rounding (~1e-6 relative, inside the 1e-4 network tolerance).
        `graph` (default: batches of >= 64 envs) replays one CUDA graph per waiting VM instead of ~40 eager launches; the
        VM counter lives on the device, so every replay is the same graph.  `fused` (default with `graph`): the replayed
        iteration is two GEMMs (both hidden heads side by side, both output heads block-diagonal) + ONE kernel for everything
        else (`vmgym_drlvmp_iter`: dueling softmax -> q -> argmax -> heuristic -> observation / feature update)."""
        vec = self.vec
        host = isinstance(observation, np.ndarray)
        obs_in = torch.from_numpy(np.ascontiguousarray(observation, np.float32)).to(self.device) if host else observation
        single = obs_in.dim() == 1
        obs_in = obs_in.reshape(-1, vec.obs_dim).float()
        n, V, P = obs_in.shape[0], vec.V, vec.P
        if graph is None:
            graph = incremental and n >= 64
        if fused is None:
            fused = graph
        fused = bool(fused and graph and self.n_actions == 4 and self.config.atom_size <= 128)
        st = self._act_state(n) if graph else None
        obs = st["obs"] if graph else obs_in.contiguous().clone()             # the working observation copy
        if graph:
            obs.copy_(obs_in)
        waiting = obs[:, :V] == float(P)                                       # fixed at entry (drlvmp.py:507)
        n_wait = waiting.sum(1)
        # k-th waiting slot of every env, in slot order
        order = torch.argsort((~waiting).to(torch.int8), dim=1, stable=True)
        rows = torch.arange(n, device=self.device)
        lin = self.dqn.feature_layer[0]
        n_iter = int(n_wait.max().item()) if n else 0

        eff = self.dqn.effective_heads() if incremental else None            # noise is fixed during act (no reset_noise)
        if graph:
            if "eff" not in st:
                st["eff"] = [(w.clone(), b.clone()) for w, b in eff]
            for (sw, sb), (w, b) in zip(st["eff"], eff):
                sw.copy_(w); sb.copy_(b)
            eff = st["eff"]

        def iteration(pre, kdev, w_cols):
            """One waiting VM per env: k-th waiting slot (k on the device), network choice, heuristic, write-back."""
            active = n_wait > kdev
            col = order.gather(1, kdev.clamp(max=V - 1).expand(n, 1)).squeeze(1)
            idx = torch.where(active, col, torch.full_like(col, -1)).to(torch.int32).contiguous()
            q = self.dqn.q_from_pre(pre, eff) if pre is not None else self.dqn(obs)
            choice = q.argmax(dim=1).to(torch.int32).contiguous()              # drlvmp.py:514-515
            pm = self.heuristic(obs, idx, choice)
            upd = active & (pm >= 0)
            # masked scatter with fixed shapes (no boolean indexing: that would synchronise with the host every VM)
            c = idx.clamp(min=0).long()
            old = obs[rows, c]
            new = torch.where(upd, pm.float(), old)
            obs[rows, c] = new                                                  # placement written into the obs copy only
            if pre is not None:
                pre += w_cols[c] * (new - old).unsqueeze(1)                     # old is WAIT (= P) wherever upd holds
            kdev += 1

        if not graph:
            kdev = torch.zeros((), dtype=torch.int64, device=self.device)
            w_cols = lin.weight.t().contiguous() if incremental else None      # [D, H]: row j = weight column of input j
            pre = None
            for k in range(n_iter):
                if incremental and k % refresh == 0:
                    pre = lin(obs)
                iteration(pre, kdev, w_cols)
        else:
            # static buffers + one captured iteration; the graph reads n_wait / order through these buffers
            st["n_wait"].copy_(n_wait); st["order"].copy_(order); st["kdev"].zero_()
            st["w_cols"].copy_(lin.weight.t())
            n_wait, order = st["n_wait"], st["order"]
            pre, kdev, w_cols = st["pre"], st["kdev"], st["w_cols"]
            if fused:
                self._act_fused(st, eff, obs, n, n_iter, refresh, lin)
            elif st["graph"] is None and n_iter > 0:
                pre.copy_(lin(obs))
                keep = obs.clone()
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    iteration(pre, kdev, w_cols)                               # warm-up (allocations) outside the capture
                torch.cuda.current_stream(self.device).wait_stream(side)
                obs.copy_(keep); kdev.zero_()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    iteration(pre, kdev, w_cols)
                st["graph"] = g
            for k in range(n_iter if not fused else 0):
                if k % refresh == 0:
                    pre.copy_(lin(obs))
                st["graph"].replay()
        action = obs[:, :V].to(torch.int64)
        if host:
            a = action.cpu().numpy()
            return a[0] if single else a
        return action[0] if single else action

    def _act_fused(self, st, eff, obs, n, n_iter, refresh, lin):
        """The per-VM iteration of act() as 2 GEMMs + `vmgym_drlvmp_iter`, captured once and replayed per waiting VM.
        The GEMMs (both hidden heads side by side, both output heads block-diagonal, and the D x H feature product of every
        `refresh`-th VM) run on the hand-written tcgen05 kernel `vmgym_tc_gemm` with split-bf16 operands ([hi | lo | hi] x
        [hi | hi | lo], ~2^-16 relative: the q-value argmax must not move), unless `self.gemm == "torch"` (cuBLAS, A/B runs)."""
        vec, dev = self.vec, self.device
        (wah, bah), (wa, ba), (wvh, bvh), (wv, bv) = eff
        Hh, H, atoms, A = wah.shape[0], wah.shape[1], self.config.atom_size, self.n_actions
        tc = getattr(self, "gemm", "tc") == "tc" and H % 8 == 0 and Hh % 4 == 0
        lib = nv.lib()
        if "WhT" not in st:
            st["WhT"] = torch.empty((H, 2 * Hh), dtype=torch.float32, device=dev)
            st["bh"] = torch.empty(2 * Hh, dtype=torch.float32, device=dev)
            ld = -(-(A + 1) * atoms // 64) * 64                                # output width padded for the GEMM's tile shapes
            st["WoT"] = torch.zeros((2 * Hh, ld), dtype=torch.float32, device=dev)
            st["bo"] = torch.zeros(ld, dtype=torch.float32, device=dev)
            st["feat"] = torch.empty((n, H), dtype=torch.float32, device=dev)
            st["graph_fused"] = None
            # tensor-core operands (bf16 splits): weights [out, 3 in] = [hi | hi | lo], activations [n, 3 in] = [hi | lo | hi]
            Dp = (vec.obs_dim + 7) // 8 * 8
            bf = torch.bfloat16
            st["Dp"] = Dp
            st["Wf_s"] = torch.zeros((H, 3 * Dp), dtype=bf, device=dev)
            st["Wh_s"] = torch.zeros((2 * Hh, 3 * H), dtype=bf, device=dev)
            st["Wo_s"] = torch.zeros((ld, 3 * 2 * Hh), dtype=bf, device=dev)
            st["obs_s"] = torch.zeros((n, 3 * Dp), dtype=bf, device=dev)
            st["feat_s"] = torch.zeros((n, 3 * H), dtype=bf, device=dev)
            st["h_s"] = torch.zeros((n, 3 * 2 * Hh), dtype=bf, device=dev)
            st["heads"] = torch.zeros((n, ld), dtype=torch.float32, device=dev)
        ld = st["bo"].shape[0]
        # both hidden heads side by side; both output heads block-diagonal (advantage atoms | value atoms)
        st["WhT"][:, :Hh].copy_(wah.t()); st["WhT"][:, Hh:].copy_(wvh.t())
        st["bh"][:Hh].copy_(bah); st["bh"][Hh:].copy_(bvh)
        st["WoT"][:Hh, :A * atoms].copy_(wa.t()); st["WoT"][Hh:, A * atoms:(A + 1) * atoms].copy_(wv.t())
        st["bo"][:A * atoms].copy_(ba); st["bo"][A * atoms:(A + 1) * atoms].copy_(bv)
        pre, feat, kdev = st["pre"], st["feat"], st["kdev"]
        ccfg = vec._ccfg()

        def stream():
            return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

        def split(src, dst, cols, cols_pad, order):
            src = src.contiguous()
            nv.check(lib.vmgym_cast_split_bf16(src.data_ptr(), src.shape[0], cols, src.stride(0), dst.data_ptr(), cols_pad, order, stream()),
                     "vmgym_cast_split_bf16")

        def gemm(a, b, M, N, K, bias, act, c32=None, c16=None):
            nv.check(lib.vmgym_tc_gemm(a.data_ptr(), 0, a.stride(0), b.data_ptr(), 0, b.stride(0), M, N, K, bias.data_ptr(), act, None, 0,
                                       c32.data_ptr() if c32 is not None else None, c32.stride(0) if c32 is not None else 0, 0,
                                       c16.data_ptr() if c16 is not None else None, c16.stride(0) if c16 is not None else 0, None, stream()),
                     "vmgym_tc_gemm")

        if tc:
            Dp = st["Dp"]
            split(lin.weight, st["Wf_s"], vec.obs_dim, Dp, 1)
            split(st["WhT"].t(), st["Wh_s"], H, H, 1)                         # [2 Hh, H] = the two hidden heads' nn.Linear weights
            split(st["WoT"].t(), st["Wo_s"], 2 * Hh, 2 * Hh, 1)

        def refresh_pre():
            """pre = obs W_f^T + b_f from scratch, feat = relu(pre) (and its bf16 split)."""
            if tc:
                split(obs, st["obs_s"], vec.obs_dim, st["Dp"], 0)
                # act = relu | fp32 output before the activation | split bf16 output
                gemm(st["obs_s"], st["Wf_s"], n, H, 3 * st["Dp"], lin.bias, 2 | 4 | 8, c32=pre, c16=st["feat_s"])
                torch.clamp(pre, min=0.0, out=feat)
            else:
                pre.copy_(lin(obs)); torch.clamp(pre, min=0.0, out=feat)

        act_mm = getattr(torch, "_addmm_activation", None)                    # addmm with the ReLU in the GEMM epilogue

        def iteration(k_offset=0):
            if tc:
                gemm(st["feat_s"], st["Wh_s"], n, 2 * Hh, 3 * H, st["bh"], 2 | 8, c16=st["h_s"])                 # relu, split out
                gemm(st["h_s"], st["Wo_s"], n, ld, 3 * 2 * Hh, st["bo"], 0, c32=st["heads"])
                heads = st["heads"]
            else:
                h = act_mm(st["bh"], feat, st["WhT"]) if act_mm is not None else torch.addmm(st["bh"], feat, st["WhT"]).relu_()
                heads = torch.addmm(st["bo"], h, st["WoT"])
            nv.check(lib.vmgym_drlvmp_iter(C.byref(ccfg), H, A, atoms, heads.data_ptr(), heads.shape[1], self.support.data_ptr(), obs.data_ptr(),
                                           st["order"].data_ptr(), st["n_wait"].data_ptr(), kdev.data_ptr(), k_offset,
                                           st["w_cols"].data_ptr(), pre.data_ptr(), feat.data_ptr(),
                                           st["feat_s"].data_ptr() if tc else None, n, stream()), "vmgym_drlvmp_iter")

        # several iterations per captured graph (fewer replays); iterations past an env's last waiting VM are no-ops
        unroll = next(u for u in (32, 16, 8, 4, 2, 1) if refresh % u == 0)       # 32: ~30 replays per act() at 1000 PMs (host-launch bound otherwise)
        if st["graph_fused"] is None:
            st["graph_fused"] = {}
        key = (unroll, tc)
        if key not in st["graph_fused"] and n_iter > 0:
            refresh_pre()
            keep = obs.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                iteration()                                                    # warm-up (allocations) outside the capture
            torch.cuda.current_stream(dev).wait_stream(side)
            obs.copy_(keep); kdev.zero_()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for u in range(unroll):
                    iteration(u)                                               # VM k = device counter + u
                kdev.add_(unroll)
            st["graph_fused"][key] = g
        for k in range(0, n_iter, unroll):
            if k % refresh == 0:
                refresh_pre()
            st["graph_fused"][key].replay()

    def _act_state(self, n: int):
        """Static buffers (and, lazily, the captured graph) of act() for a batch of n observations."""
        cache = self.__dict__.setdefault("_act_cache", {})
        st = cache.get(n)
        if st is None:
            vec, dev = self.vec, self.device
            H = self.dqn.feature_layer[0].out_features
            st = dict(obs=torch.empty((n, vec.obs_dim), dtype=torch.float32, device=dev),
                      pre=torch.empty((n, H), dtype=torch.float32, device=dev),
                      n_wait=torch.empty(n, dtype=torch.int64, device=dev),
                      order=torch.empty((n, vec.V), dtype=torch.int64, device=dev),
                      kdev=torch.zeros((), dtype=torch.int64, device=dev),
                      w_cols=torch.empty((vec.obs_dim, H), dtype=torch.float32, device=dev), graph=None)
            cache[n] = st
        return st


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
