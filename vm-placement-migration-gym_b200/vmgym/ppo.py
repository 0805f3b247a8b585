"""PPO agent of the reference (src/agents/ppo.py) on the batched device env.

Same interface — `PPOConfig` fields (ppo.py:14-36), `Network` (actor / critic 3-layer tanh MLPs with the reference's
orthogonal init, ppo.py:85-109), `PPOAgent.act / learn / update / save_model / load_model / eval` — with the hot numeric
pieces on hand-written kernels:

  * masked multi-categorical heads (mask built on the fly from the env records, migration-ratio gating, sampling,
    log-prob, entropy) and their backward: `vmgym_policy_heads{,_backward}` (csrc/vmgym_policy.cu);
  * GAE: `vmgym_gae` (warp-level reverse scan);
  * env stepping / masks: the env kernels.
The dense layers run through torch (cuBLAS) in this round; autograd, AdamW and the NCCL gradient all-reduce are
torch plumbing.  Sampling uses a Philox stream, so sampled trajectories differ from torch's CPU generator by
construction; logits / log-prob / entropy / GAE are parity-tested against the reference formulas.
"""
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
class PPOConfig:
    episodes: int = 2000
    hidden_size: int = 256
    migration_ratio: float = 0.5
    masked: bool = True
    lr: float = 5e-5
    gamma: float = 0.99
    lamda: float = 0.98
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    vf_loss_clip: bool = True
    k_epochs: int = 4
    kl_max: float = 0.02
    eps_clip: float = 0.1
    max_grad_norm: float = 0.5
    batch_size: int = 100          # rollout length T (steps) between updates
    minibatch_size: int = 25       # time steps per minibatch (sequential, ppo.py:251-252)
    det: bool = False
    network_arch: str = "separate"
    reward_scaling: bool = False
    training_progress_bar: bool = True
    device: str = "cuda"
    env_chunk: int = 2048          # samples (env-steps) per forward/backward chunk inside a minibatch (gradient accumulation)
    fused_rollout: bool = False    # rollouts through the fused tcgen05 actor head (bf16 operands; action_dim <= 128)
    update_math: str = "auto"      # "bf16": forward + backward of update() and the rollout forward on the hand-written tcgen05 kernels
                                   # (vmgym/ppo_tc.py; needs action_dim <= 128, hidden % 64 == 0, hidden <= 512); "fp32": torch
                                   # autograd on fp32 layers — the parity mode against the reference; "auto": bf16 where supported
    vf_broadcast: bool = False     # reference-exact value loss: ppo.py:274-277 subtracts returns [mb] from newvalues [mb, 1], which
                                   # broadcasts to [mb, mb] (every value against every return); False = the elementwise loss it
                                   # evidently means (DESIGN.md §2 "intentional deviations").  O(mb^2) memory: small batches only


def _ortho(layer: nn.Linear, gain: float) -> nn.Linear:
    nn.init.orthogonal_(layer.weight, gain=gain)       # ppo.py:85-88
    nn.init.constant_(layer.bias, 0.0)
    return layer


class Network(nn.Module):
    """ppo.py:91-131: separate critic (-> 1) and actor (-> sum(nvec) logits), tanh, orthogonal init (gains sqrt2,
    sqrt2, then 1 / 0.01)."""

    def __init__(self, input_size: int, n_vms: int, action_dim: int, hidden_size: int, dtype=torch.float32):
        super().__init__()
        self.n_vms, self.action_dim = int(n_vms), int(action_dim)
        g = math.sqrt(2.0)
        self.critic = nn.Sequential(_ortho(nn.Linear(input_size, hidden_size, dtype=dtype), g), nn.Tanh(),
                                    _ortho(nn.Linear(hidden_size, hidden_size, dtype=dtype), g), nn.Tanh(),
                                    _ortho(nn.Linear(hidden_size, 1, dtype=dtype), 1.0))
        self.actor = nn.Sequential(_ortho(nn.Linear(input_size, hidden_size, dtype=dtype), g), nn.Tanh(),
                                   _ortho(nn.Linear(hidden_size, hidden_size, dtype=dtype), g), nn.Tanh(),
                                   _ortho(nn.Linear(hidden_size, self.n_vms * self.action_dim, dtype=dtype), 0.01))

    def get_value(self, obs):
        return self.critic(obs)

    def get_det_action(self, obs):
        """ppo.py:128-131: per-VM argmax of the UNMASKED logits."""
        return self.actor(obs).reshape(-1, self.n_vms, self.action_dim).argmax(dim=-1)


class _MaskedHeads(torch.autograd.Function):
    """(logits, stored packed mask, stored actions) -> (sum log-prob, sum entropy) per env, with the backward kernel."""

    @staticmethod
    def forward(ctx, logits, mask_bits, actions, ccfg, masked):
        logits = logits.contiguous()
        n = logits.shape[0]
        logprob = torch.empty(n, dtype=torch.float32, device=logits.device)
        entropy = torch.empty(n, dtype=torch.float32, device=logits.device)
        code = {torch.uint8: nv.U8, torch.int16: nv.I16, torch.int64: nv.I64}[actions.dtype]
        stream = C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
        nv.check(nv.lib().vmgym_policy_heads(C.byref(ccfg), None, mask_bits.data_ptr() if mask_bits is not None else None,
                                             int(masked), logits.data_ptr(), n, actions.data_ptr(), code, -1.0, 0, 0, None,
                                             logprob.data_ptr(), entropy.data_ptr(), None, stream), "vmgym_policy_heads")
        ctx.save_for_backward(logits, mask_bits if mask_bits is not None else torch.empty(0, device=logits.device), actions)
        ctx.ccfg, ctx.masked, ctx.code, ctx.has_mask = ccfg, masked, code, mask_bits is not None
        return logprob, entropy

    @staticmethod
    def backward(ctx, g_logprob, g_entropy):
        logits, mask_bits, actions = ctx.saved_tensors
        g_logits = torch.empty_like(logits)
        n = logits.shape[0]
        stream = C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
        g_logprob = g_logprob.contiguous().float()
        g_entropy = g_entropy.contiguous().float()
        nv.check(nv.lib().vmgym_policy_heads_backward(C.byref(ctx.ccfg), mask_bits.data_ptr() if ctx.has_mask else None,
                                                      int(ctx.masked), logits.data_ptr(), n, actions.data_ptr(), ctx.code,
                                                      g_logprob.data_ptr(), g_entropy.data_ptr(), g_logits.data_ptr(), stream),
                 "vmgym_policy_heads_backward")
        return g_logits, None, None, None, None


class _OutLinear(torch.autograd.Function):
    """The actor's output layer `h @ W^T + b` (ppo.py:103-109) for the update.  Forward is `F.linear`; the backward computes the
    bias gradient inside the weight-gradient GEMM (a ones column appended to `h`), instead of a separate column-sum pass over
    the [samples, V*A] logit gradients (4 GB per 32768-sample minibatch at config/100.yml: 8 % of an update's device time)."""

    PAD = 16          # columns appended to h: [1, 0, ..., 0] (keeps the GEMM's N a multiple of 16)

    @staticmethod
    def forward(ctx, h, weight, bias):
        ctx.save_for_backward(h, weight)
        return F.linear(h, weight, bias)

    @staticmethod
    def backward(ctx, g):
        h, weight = ctx.saved_tensors
        g = g.contiguous()
        H = h.shape[1]
        h_aug = torch.zeros((h.shape[0], H + _OutLinear.PAD), dtype=h.dtype, device=h.device)
        h_aug[:, :H] = h
        h_aug[:, H] = 1.0
        gh = g @ weight if ctx.needs_input_grad[0] else None
        gwb = g.t() @ h_aug                                   # [V*A, H + PAD]: weight gradient | bias gradient | zeros
        return gh, gwb[:, :H].contiguous(), gwb[:, H].contiguous()


def linear_bf16(a: torch.Tensor, weight_bf16: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """a[M, K] @ weight[N, K]^T + bias on the hand-written tcgen05 kernel (bf16 operands, fp32 accumulate / output)."""
    a = a.to(torch.bfloat16).contiguous()
    M, K = a.shape
    N = weight_bf16.shape[0]
    assert weight_bf16.dtype == torch.bfloat16 and weight_bf16.is_contiguous() and weight_bf16.shape[1] == K
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    b = bias.float().contiguous() if bias is not None else None
    stream = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
    nv.check(nv.lib().vmgym_linear_bf16(a.data_ptr(), weight_bf16.data_ptr(), b.data_ptr() if b is not None else None,
                                        out.data_ptr(), M, N, K, N, stream), "vmgym_linear_bf16")
    return out


class FusedActorHead:
    """The actor's output layer + masked multi-categorical heads as ONE tcgen05 kernel (vmgym_policy_fused): logits live
    only in tensor memory.  Holds the bf16 re-layout of nn.Linear(hidden, V*A): VM v's A rows at [R v, R v + A), R = TILE."""

    def __init__(self, linear: nn.Linear, n_vms: int, action_dim: int):
        if action_dim > 128:
            raise nv.VmgymError("the fused actor head needs action_dim <= 128")
        self.V, self.A, self.K = int(n_vms), int(action_dim), linear.in_features
        self.TILE = int(nv.lib().vmgym_policy_fused_rows(self.A, self.K))     # rows per VM of the padded operands (include/vmgym.h)
        self.linear = linear
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """Re-derive the padded bf16 weights after an optimiser step."""
        V, A, K, T = self.V, self.A, self.K, self.TILE
        w = torch.zeros((V, T, K), dtype=torch.bfloat16, device=self.linear.weight.device)
        w[:, :A] = self.linear.weight.detach().view(V, A, K).to(torch.bfloat16)
        b = torch.zeros((V, T), dtype=torch.float32, device=w.device)
        b[:, :A] = self.linear.bias.detach().view(V, A).float()
        self.w_pad, self.b_pad = w.view(V * T, K).contiguous(), b.view(-1).contiguous()

    def __call__(self, hidden, mask_bits, seed: int, counter: int, action_in=None):
        """hidden [M, K] -> (action u8 [M, V], logprob [M], entropy [M]); `mask_bits` int32 [M, V, 4] or None."""
        h = hidden.to(torch.bfloat16).contiguous()
        M = h.shape[0]
        dev = h.device
        action = torch.empty((M, self.V), dtype=torch.uint8, device=dev) if action_in is None else None
        lp = torch.empty((M, self.V), dtype=torch.float32, device=dev)
        ent = torch.empty((M, self.V), dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        nv.check(nv.lib().vmgym_policy_fused(h.data_ptr(), self.w_pad.data_ptr(), self.b_pad.data_ptr(),
                                             mask_bits.data_ptr() if mask_bits is not None else None,
                                             action_in.data_ptr() if action_in is not None else None, M, self.V, self.A, self.K,
                                             int(seed), int(counter), action.data_ptr() if action is not None else None,
                                             lp.data_ptr(), ent.data_ptr(), stream), "vmgym_policy_fused")
        return (action if action_in is None else action_in), lp.sum(1), ent.sum(1)


def gae(rewards, values, next_values, dones, gamma: float, lamda: float):
    """ppo.py:237-243 on time-major [T, N] float32 tensors (dones uint8/bool).  Returns (advantages, returns)."""
    T, N = rewards.shape
    rewards, values, next_values = rewards.contiguous().float(), values.contiguous().float(), next_values.contiguous().float()
    dones = dones.contiguous().to(torch.uint8)
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    stream = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    nv.check(nv.lib().vmgym_gae(rewards.data_ptr(), values.data_ptr(), next_values.data_ptr(), dones.data_ptr(), T, N,
                                float(gamma), float(lamda), adv.data_ptr(), ret.data_ptr(), stream), "vmgym_gae")
    return adv, ret


class PPOAgent(AgentBase):
    """Reference-shaped agent (Base API: learn / act / save_model / load_model / eval) over a VecVmEnv or VmEnv."""

    name = "PPOAgent"

    def __init__(self, env, config: PPOConfig | None = None, seed: int | None = None):
        self.env = env
        self.vec = getattr(env, "vec", env)
        self.config = config or PPOConfig()
        vec = self.vec
        self.device = vec.device
        self.obs_dim, self.V, self.A = vec.obs_dim, vec.V, vec.action_dim
        self.model = Network(self.obs_dim, self.V, self.A, self.config.hidden_size).to(self.device)
        self._flatten_parameters()
        self.mask_words = (self.A + 31) // 32
        self.seed = int(vec.config.seed if seed is None else seed)
        self._calls = 0
        self._fused = None
        self.total_steps = 0
        self.training = True
        self.data_parallel = True      # use the initialised torch.distributed group (one rank per GPU) in update() / episode_seeds()

    # ---- optimiser state: AdamW(lr) of ppo.py:143 on flat buffers (vmgym_adamw_step) ----------------------------
    def _flatten_parameters(self):
        """Every parameter becomes a view of ONE flat fp32 buffer and every gradient a view of a second one, so the
        data-parallel all-reduce, the gradient-norm clip and the AdamW step each run once over contiguous memory (no
        concatenate / copy-back).  load_state_dict copies in place, so the views survive it."""
        params = list(self.model.parameters())
        ALIGN = 64                       # floats: every parameter starts on a 256-byte boundary (GEMM / TMA operand alignment);
        offs, n_pad = [], 0              # the gaps hold zeros, which AdamW leaves at zero
        for p in params:
            offs.append(n_pad)
            n_pad += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        dev = self.device
        self._flat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self._flat_grad = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        for p, off in zip(params, offs):
            k = p.numel()
            self._flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self._flat[off:off + k].view_as(p)
            p.grad = self._flat_grad[off:off + k].view_as(p)
        self._n_params = n_pad
        self._exp_avg = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self._exp_avg_sq = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self._opt_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self._opt_ws = torch.zeros(1024, dtype=torch.float64, device=dev)
        self._opt_skip = torch.zeros(1, dtype=torch.int32, device=dev)           # sticky within an epoch (KL early stop)
        self._grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.adam = dict(beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01)    # torch.optim.AdamW defaults (ppo.py:143)

    def _tc_network(self):
        """The tensor-core network (vmgym/ppo_tc.py) when config.update_math selects it and the shape allows it, else None."""
        mode = self.config.update_math
        if mode not in ("auto", "bf16", "fp32"):
            raise ValueError("update_math must be 'auto', 'bf16' or 'fp32'")
        H = self.config.hidden_size
        ok = self.A <= 128 and H <= 512 and H % 64 == 0 and not self.config.vf_broadcast and self.vec.place_dtype == torch.uint8
        if mode == "fp32" or (mode == "auto" and not ok):
            return None
        if not ok:
            raise nv.VmgymError("update_math='bf16' needs action_dim <= 128, hidden % 64 == 0, hidden <= 512 and vf_broadcast off")
        if getattr(self, "_tc", None) is None:
            from .ppo_tc import TensorCoreNetwork
            self._tc = TensorCoreNetwork(self, self.config.env_chunk)
        return self._tc

    def _mask4(self, mask):
        """Packed mask rows padded to the 4 words the fused tensor-core head reads per (env, VM)."""
        if mask is None or mask.shape[-1] == 4:
            return mask
        out = torch.zeros(mask.shape[:-1] + (4,), dtype=torch.int32, device=mask.device)
        out[..., :mask.shape[-1]] = mask
        return out

    def _optim_step(self, grad_scale: float = 1.0):
        """clip_grad_norm_(max_grad_norm) + AdamW step (ppo.py:284-287) unless the device-side skip flag is set."""
        cfg, a = self.config, self.adam
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(nv.lib().vmgym_adamw_step(self._flat.data_ptr(), self._flat_grad.data_ptr(), self._exp_avg.data_ptr(),
                                           self._exp_avg_sq.data_ptr(), self._n_params, float(cfg.lr), a["beta1"], a["beta2"], a["eps"],
                                           a["weight_decay"], float(cfg.max_grad_norm or 0.0), float(grad_scale),
                                           self._opt_ws.data_ptr(), self._opt_skip.data_ptr(), self._opt_step.data_ptr(),
                                           self._grad_norm.data_ptr(), stream), "vmgym_adamw_step")

    # ---- reference API -------------------------------------------------------------------------------------
    def eval(self, mode=True):
        self.training = not mode
        self.model.train(not mode)

    def save_model(self, modelpath):
        if modelpath:
            # the reference saves the torch.compile-wrapped module: keys carry the `_orig_mod.` prefix (ppo.py:142,166)
            torch.save({"_orig_mod." + k: v for k, v in self.model.state_dict().items()}, modelpath)

    def load_model(self, modelpath):
        sd = torch.load(modelpath, map_location=self.device)
        sd = {k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k: v for k, v in sd.items()}
        self.model.load_state_dict(sd)
        self.model.eval()
        self.weights_changed()

    def weights_changed(self):
        """Call after writing the parameters from outside (load_state_dict, broadcast): re-derives the bf16 operand copies."""
        if getattr(self, "_tc", None) is not None:
            self._tc.refresh()
        elif self._fused is not None:
            self._fused.refresh()

    def _heads(self, logits, migration_ratio: float, want_mask: bool):
        vec = self.vec
        n = logits.shape[0]
        action = torch.empty((n, self.V), dtype=vec.place_dtype, device=self.device)
        logprob = torch.empty(n, dtype=torch.float32, device=self.device)
        entropy = torch.empty(n, dtype=torch.float32, device=self.device)
        mask = torch.empty((n, self.V, self.mask_words), dtype=torch.int32, device=self.device) if want_mask else None
        self._calls += 1
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(nv.lib().vmgym_policy_heads(C.byref(vec._ccfg()), vec.state.data_ptr(), None, int(bool(self.config.masked)),
                                             logits.data_ptr(), n, None, nv.U8, float(migration_ratio), self.seed, self._calls,
                                             action.data_ptr(), logprob.data_ptr(), entropy.data_ptr(),
                                             mask.data_ptr() if mask is not None else None, stream), "vmgym_policy_heads")
        return action, logprob, entropy, mask

    def _mask_bits(self, migration_ratio: float):
        """Packed invalid-action bits [N, V, ceil(A/32)] of the env's current state (+ gating), no logits involved."""
        vec = self.vec
        n = vec.num_envs
        mask = torch.empty((n, self.V, self.mask_words), dtype=torch.int32, device=self.device)
        self._calls += 1
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(nv.lib().vmgym_policy_heads(C.byref(vec._ccfg()), vec.state.data_ptr(), None, int(bool(self.config.masked)),
                                             None, n, None, nv.U8, float(migration_ratio), self.seed, self._calls, None, None,
                                             None, mask.data_ptr(), stream), "vmgym_policy_heads(mask)")
        return mask

    @torch.no_grad()
    def fused_sample(self, obs, migration_ratio: float = -1.0):
        """Rollout forward on tensor cores: obs -> hidden (torch) -> fused output layer + heads (tcgen05, bf16 operands).
        Returns (action u8 [N, V], logprob [N], entropy [N], packed mask).  Needs action_dim <= 128."""
        tc = self._tc_network()
        if self._fused is None:
            self._fused = FusedActorHead(self.model.actor[4], self.V, self.A)
        hidden = tc.actor_hidden(obs) if tc is not None else self.model.actor[:4](obs)
        mask = self._mask_bits(migration_ratio)
        self._calls += 1
        action, lp, ent = self._fused(hidden, self._mask4(mask) if self.config.masked else None, self.seed, self._calls)
        return action, lp, ent, mask

    @torch.no_grad()
    def act(self, obs):
        """ppo.py:151-161: mask + migration-ratio gating + sampled (or deterministic) action for the env's CURRENT state."""
        single = False
        if isinstance(obs, np.ndarray):
            single = obs.ndim == 1
            obs_t = torch.from_numpy(np.ascontiguousarray(obs, np.float32).reshape(-1, self.obs_dim)).to(self.device)
        else:
            single = obs.dim() == 1
            obs_t = obs.reshape(-1, self.obs_dim).to(self.device, torch.float32)
        if self.config.det:
            action = self.model.get_det_action(obs_t).to(self.vec.place_dtype)
        else:
            logits = self.model.actor(obs_t).contiguous()
            action, _, _, _ = self._heads(logits, self.config.migration_ratio, want_mask=False)
        if isinstance(obs, np.ndarray):
            a = action.cpu().numpy().astype(np.int64)
            return a[0] if single else a
        return action[0] if single else action

    @torch.no_grad()
    def rollout(self, n_steps: int, fused: bool | None = None):
        """Evaluation rollout (Base.test with PPOAgent.act, base.py:71-86 + ppo.py:151-161) on all envs in lock-step:
        mask + migration-ratio gating + sampled action + env.step, `n_steps` times from the envs' current state.
        `fused` (default: whenever the action space allows it) uses the tcgen05 fused actor head.  Returns the per-env
        sum of rewards (float64, device)."""
        vec = self.vec
        use_fused = (self.mask_words == 4 and self.A <= 128) if fused is None else fused
        obs = vec.observe()
        ret = torch.zeros(vec.num_envs, dtype=torch.float64, device=self.device)
        for _ in range(int(n_steps)):
            if use_fused and not self.config.det:
                action = self.fused_sample(obs, self.config.migration_ratio)[0]
            else:
                action = self.act(obs)
            obs, reward, term, _, _ = vec.step(action, want_valid=False)
            ret += reward
        return ret

    # ---- training ------------------------------------------------------------------------------------------
    def episode_seeds(self, ep: int):
        """Seeds of episode `ep`.  One env: `config.seed + ep`, the reference's (ppo.py:192).  A batch: env g (global id =
        construction seed - config.seed, so shards of a multi-GPU job stay distinct) gets `config.seed + 4 (ep * n_global + g)`:
        the reference's four generators sit at seed .. seed + 3 (env.py:175-178), so a stride of 4 keeps every stream of every
        env and episode disjoint (consecutive seeds would replay env i+1's arrival trace in env i one episode later, and alias
        env i's memory-size stream with env i+1's cpu-size stream)."""
        vec = self.vec
        base = int(vec.config.seed)
        world = torch.distributed.get_world_size() if self.data_parallel and torch.distributed.is_available() and \
            torch.distributed.is_initialized() else 1
        n_global = vec.num_envs * world
        if n_global == 1:
            return np.asarray([base + ep], np.int64)
        gid = getattr(self, "_env_ids", None)
        if gid is None:
            s0 = np.asarray(vec._seeds, np.int64) if vec._seeds is not None else base + np.arange(vec.num_envs, dtype=np.int64)
            gid = self._env_ids = s0 - base
        return base + 4 * (ep * n_global + gid)

    def learn(self, episodes: int | None = None, max_updates: int | None = None, reset: bool = True):
        """ppo.py:172-226 with N envs stepping in lock-step: every `batch_size` steps one `update` on the [T, N] rollout; the
        rollout cursor and buffers persist across episodes like the reference's `i_batch` (ppo.py:181,209-214).
        Runs `episodes` episodes of `training_steps` steps (the reference's loop bound is a known slip).
        `reset=False` continues from the envs' current state instead of resetting at the start of each episode (benchmarks)."""
        cfg, vec = self.config, self.vec
        T, N = cfg.batch_size, vec.num_envs
        dev = self.device
        buf = getattr(self, "_rollout", None)
        if buf is None or buf["reward"].shape != (T, N):
            buf = self._rollout = dict(
                obs=torch.empty((T, N, self.obs_dim), dtype=torch.float32, device=dev),
                next_obs=torch.empty((T, N, self.obs_dim), dtype=torch.float32, device=dev),
                action=torch.empty((T, N, self.V), dtype=vec.place_dtype, device=dev),
                mask=torch.empty((T, N, self.V, self.mask_words), dtype=torch.int32, device=dev),
                logprob=torch.empty((T, N), dtype=torch.float32, device=dev),
                reward=torch.empty((T, N), dtype=torch.float32, device=dev),
                done=torch.empty((T, N), dtype=torch.uint8, device=dev))
            self._cursor = 0
        returns, updates = [], 0
        vec.eval(False)
        for ep in range(int(cfg.episodes if episodes is None else episodes)):
            if reset:
                obs, _ = vec.reset(seed=self.episode_seeds(ep))      # ppo.py:192
            else:
                obs = vec.observe()
            obs = obs.clone()
            ep_ret = torch.zeros(N, dtype=torch.float64, device=dev)
            done = False
            while not done:
                i = self._cursor
                with torch.no_grad():
                    if cfg.fused_rollout:
                        action, logprob, _, mask = self.fused_sample(obs, -1.0)
                    else:
                        logits = self.model.actor(obs).contiguous()
                        action, logprob, _, mask = self._heads(logits, -1.0, want_mask=True)   # no gating in training (ppo.py:196-197)
                nobs, reward, term, _, _ = vec.step(action, want_valid=False)
                buf["obs"][i], buf["next_obs"][i], buf["action"][i], buf["mask"][i] = obs, nobs, action, mask
                buf["logprob"][i], buf["reward"][i], buf["done"][i] = logprob, reward.float(), vec.terminated_u8
                ep_ret += reward
                obs = nobs.clone()
                self._cursor += 1
                self.total_steps += 1
                done = bool(term[0].item())                           # all envs share the step limit
                if self._cursor >= T:
                    self.update(**buf)
                    if self._fused is not None:
                        self._fused.refresh()
                    self._cursor = 0
                    updates += 1
                    if max_updates is not None and updates >= max_updates:
                        return returns
            returns.append(ep_ret.mean().item())
        return returns

    def _minibatch_backward(self, obs_mb, act_mb, mask_mb, lp_mb, adv_mb, val_mb, ret_mb, n_total: int):
        """Forward + backward of one minibatch (ppo.py:255-285) in chunks of `env_chunk` samples, gradients accumulated into the
        flat gradient buffer.  `adv_mb` is already normalised; `n_total` = samples of the minibatch over ALL ranks (the means of
        ppo.py:269,277,282 are over the global minibatch).  Returns (sum of log-ratios, loss) as device scalars (this rank's part)."""
        cfg = self.config
        ccfg = self.vec._ccfg()
        n_mb = obs_mb.shape[0]
        self._flat_grad.zero_()
        tc = self._tc_network()
        if tc is not None:
            # hand-written forward + backward (vmgym/ppo_tc.py); obs_mb is the bf16 operand cache built once per update
            tc.begin_minibatch()
            for s0 in range(0, n_mb, cfg.env_chunk):
                s1 = min(n_mb, s0 + cfg.env_chunk)
                tc.forward_backward(obs_mb[s0:s1], mask_mb[s0:s1] if mask_mb is not None else None, act_mb[s0:s1], lp_mb[s0:s1],
                                    adv_mb[s0:s1].contiguous(), val_mb[s0:s1], ret_mb[s0:s1], n_total)
            return tc.end_minibatch()
        logratio_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        loss_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        if cfg.vf_broadcast and n_mb > cfg.env_chunk:
            raise nv.VmgymError("vf_broadcast needs the whole minibatch in one chunk (samples <= env_chunk)")
        out = self.model.actor[4]
        for s0 in range(0, n_mb, cfg.env_chunk):
            s1 = min(n_mb, s0 + cfg.env_chunk)
            o = obs_mb[s0:s1]
            lg = _OutLinear.apply(self.model.actor[:4](o), out.weight, out.bias)
            nlp, ent = _MaskedHeads.apply(lg, mask_mb[s0:s1], act_mb[s0:s1], ccfg, cfg.masked)
            logratio = nlp - lp_mb[s0:s1]
            logratio_sum += logratio.detach().double().sum()
            ratios = torch.exp(logratio)
            a = adv_mb[s0:s1]
            loss_clipped = torch.max(-ratios * a, -torch.clamp(ratios, 1 - cfg.eps_clip, 1 + cfg.eps_clip) * a).sum()
            newv = self.model.get_value(o).flatten()
            v_old, ret = val_mb[s0:s1], ret_mb[s0:s1]
            if cfg.vf_broadcast:
                # ppo.py:272-277 as written: newvalues is [mb, 1], returns[minibatch] is [mb] -> [mb, mb] pairs, mean;
                # (i, j) pairs new value i against old value j / return j
                nv2 = newv[:, None]
                l_un = torch.square(nv2 - ret)
                l_cl = torch.square(v_old + torch.clamp(nv2 - v_old, -cfg.eps_clip, cfg.eps_clip) - ret)
                loss_vf = 0.5 * (torch.max(l_un, l_cl) if cfg.vf_loss_clip else l_un).sum() / n_mb
            else:
                l_un = torch.square(newv - ret)
                l_cl = torch.square(v_old + torch.clamp(newv - v_old, -cfg.eps_clip, cfg.eps_clip) - ret)
                loss_vf = 0.5 * (torch.max(l_un, l_cl) if cfg.vf_loss_clip else l_un).sum()
            loss = (loss_clipped - cfg.ent_coef * ent.sum() + cfg.vf_coef * loss_vf) / n_total
            loss_sum += loss.detach().double()
            loss.backward()
        return logratio_sum, loss_sum

    def _normalise_advantages(self, adv, world: int):
        """ppo.py:255-256 over the GLOBAL minibatch: (adv - mean) / (unbiased std + 1e-10).  With several ranks the three
        sums (n, sum, sum of squares) are all-reduced, so k GPUs x N/k envs normalise exactly like one GPU x N envs."""
        if world == 1:
            return ((adv - adv.mean()) / (adv.std() + 1e-10)).reshape(-1)
        a64 = adv.double()
        st = torch.stack([torch.tensor(float(adv.numel()), dtype=torch.float64, device=adv.device), a64.sum(), (a64 * a64).sum()])
        torch.distributed.all_reduce(st)
        n, mean = st[0], st[1] / st[0]
        var = (st[2] - n * mean * mean) / (n - 1.0)
        return ((a64 - mean) / (var.clamp_min(0.0).sqrt() + 1e-10)).float().reshape(-1)

    def update(self, obs, next_obs, action, mask, logprob, reward, done, debug: bool = False):
        """ppo.py:229-295 on a time-major rollout [T, N, ...].  The minibatch loop enqueues without host synchronisation: the KL
        early stop (ppo.py:263-264) is a device-side flag that turns the optimiser step of the offending minibatch — and of the
        rest of its epoch — into a no-op; the host reads the flag one minibatch late (it is already known by then) and stops
        enqueuing the epoch.  `debug`: also return what the update computed on the way (values, advantages, returns, and per
        attempted minibatch: KL, loss, pre-clip gradient norm, stepped) — with a host synchronisation per minibatch."""
        cfg = self.config
        T, N = reward.shape
        dist = torch.distributed
        world = dist.get_world_size() if self.data_parallel and dist.is_available() and dist.is_initialized() else 1
        tc = self._tc_network()
        with torch.no_grad():
            if tc is not None:
                values = tc.values(obs).reshape(T, N)
                next_values = tc.values(next_obs).reshape(T, N)
                obs = tc.cast_obs(obs).reshape(T, N, tc.Dx)             # bf16 operand of every minibatch pass, cast once
                mask = self._mask4(mask)
            else:
                values = self.model.get_value(obs.reshape(T * N, -1)).reshape(T, N)
                next_values = self.model.get_value(next_obs.reshape(T * N, -1)).reshape(T, N)
            advantages, returns = gae(reward, values, next_values, done, cfg.gamma, cfg.lamda)
        stats = {}
        attempts = []
        D = obs.shape[-1]
        flag_host = getattr(self, "_skip_host", None)
        if flag_host is None and not debug:
            flag_host = self._skip_host = torch.zeros(2, dtype=torch.int32).pin_memory()      # ping-pong: minibatch j -> slot j & 1
            self._skip_events = [torch.cuda.Event(), torch.cuda.Event()]
        kl_dev = torch.zeros((), dtype=torch.float64, device=self.device)
        for epoch in range(cfg.k_epochs):
            self._opt_skip.zero_()
            for j, t0 in enumerate(range(0, T, cfg.minibatch_size)):   # sequential minibatches (ppo.py:251-252)
                if not debug and cfg.kl_max is not None and j >= 2:
                    # the flag of minibatch j - 2 (minibatch j - 1 is still queued, so the device never runs dry on this wait)
                    self._skip_events[j & 1].synchronize()
                    if int(flag_host[j & 1]) != 0:
                        break                                          # the KL stop fired: the rest of the epoch would be no-ops
                t1 = min(T, t0 + cfg.minibatch_size)
                n_mb = (t1 - t0) * N
                adv_mb = self._normalise_advantages(advantages[t0:t1], world)
                # the minibatch as one flat list of samples (time-major rollout -> these are views)
                obs_mb, act_mb = obs[t0:t1].reshape(n_mb, D), action[t0:t1].reshape(n_mb, self.V)
                mask_mb = mask[t0:t1].reshape(n_mb, self.V, mask.shape[-1])
                lp_mb, val_mb, ret_mb = logprob[t0:t1].reshape(-1), values[t0:t1].reshape(-1), returns[t0:t1].reshape(-1)
                logratio_sum, loss_sum = self._minibatch_backward(obs_mb, act_mb, mask_mb, lp_mb, adv_mb, val_mb, ret_mb, n_mb * world)
                # KL early stop on the whole (global) minibatch (ppo.py:263-264): the reference breaks before its backward;
                # here the step is skipped on the device instead (same parameters afterwards)
                if world > 1:
                    red = torch.stack([logratio_sum, loss_sum])
                    dist.all_reduce(red)
                    logratio_sum, loss_sum = red[0], red[1]
                    dist.all_reduce(self._flat_grad)                   # data parallel: sum of the ranks' gradient parts (NCCL), in place
                kl_dev = -(logratio_sum / (n_mb * world))
                if cfg.kl_max is not None:
                    self._opt_skip.logical_or_(kl_dev > cfg.kl_max)    # sticky until the end of the epoch
                self._optim_step()
                if tc is not None:
                    tc.refresh()
                if debug:
                    torch.cuda.synchronize(self.device)
                    skipped = int(self._opt_skip.item()) != 0
                    attempts.append(dict(epoch=epoch, mb=t0 // cfg.minibatch_size, kl=float(kl_dev.item()), loss=float(loss_sum.item()),
                                         stepped=0 if skipped else 1, grad_norm=float(self._grad_norm.item())))
                    if skipped:
                        break
                elif cfg.kl_max is not None:
                    flag_host[j & 1:(j & 1) + 1].copy_(self._opt_skip, non_blocking=True)
                    self._skip_events[j & 1].record(torch.cuda.current_stream(self.device))
        stats["kl"] = kl_dev                                           # device scalar of the last minibatch (no sync here)
        if debug:
            stats.update(kl=float(kl_dev.item()), values=values, next_values=next_values, advantages=advantages, returns=returns,
                         attempts=attempts)
        return stats
