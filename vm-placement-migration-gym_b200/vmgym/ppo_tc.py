"""The PPO networks (src/agents/ppo.py:91-109) on hand-written tcgen05 kernels: forward AND backward of the actor and the
critic for PPOAgent.update (ppo.py:229-295) and the rollout forward, bf16 operands / fp32 accumulation.  No autograd, no
cuBLAS: every dense contraction is `vmgym_tc_gemm` (csrc/vmgym_train.cu) or the fused actor head (csrc/vmgym_gemm.cu).

Per chunk of samples (x = bf16 observations [M, Dp], zero padded to a multiple of 8 columns):
    actor   a1 = tanh(x W1^T + b1), a2 = tanh(a1 W2^T + b2)                      two GEMMs, tanh + bf16 in the epilogue
            (logprob, entropy)[M, V] = fused head(a2, W3, b3, mask, actions)     logits only in tensor memory
    critic  c1, c2 likewise, value = c2 w3 + b3                                  value-head kernel
    loss    clipped surrogate / clipped value loss / entropy bonus               vmgym_ppo_loss -> dL/dlogprob, dL/dvalue per sample
    back    g = dL/dlogits (bf16 [M, 128 V])                                     fused head, gradient mode (logits recomputed)
            dW3 += g^T a2, db3 += sum g ;  dz2 = (g W3) * (1 - a2^2)             GEMMs: MN x MN (+ row sums), K x MN (+ tanh backward)
            dW2 += dz2^T a1, db2 ;  dz1 = (dz2 W2) * (1 - a1^2) ;  dW1 += dz1^T x, db1
            critic: dzc2 = dv w3 (1 - c2^2), dw3, db3 (value-head backward), then the same two layers
Weight gradients land directly in the agent's flat fp32 gradient buffer (the output layer's through a 128-row-per-VM padded
scratch that is folded back once per minibatch); weights are re-cast to bf16 after every optimiser step.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nv


def _ptr(t):
    return t.data_ptr() if t is not None else None


class TensorCoreNetwork:
    def __init__(self, agent, max_chunk: int):
        from .ppo import FusedActorHead
        self.agent = agent
        m = agent.model
        self.dev = agent.device
        self.V, self.A, self.D = agent.V, agent.A, agent.obs_dim
        self.H = m.actor[0].out_features
        if self.A > 128 or self.H > 512 or self.H % 64:
            raise nv.VmgymError("the tensor-core PPO path needs action_dim <= 128 and a hidden size that is a multiple of 64, <= 512")
        self.Dp = (self.D + 7) // 8 * 8
        if agent._fused is None:
            agent._fused = FusedActorHead(m.actor[4], self.V, self.A)
        self.head = agent._fused
        self.TILE = self.head.TILE       # rows of the padded output layer per VM (FusedActorHead layout)
        H, Dp, dev = self.H, self.Dp, self.dev
        bf = torch.bfloat16
        # first layer: split operands [hi | hi | lo] x [hi | lo | hi] over a 3 Dp-wide K (vmgym_cast_split_bf16): raw observations mix
        # PM indices up to P + 1 with sizes in [0, 1], and plain bf16 operands move the pre-activations by ~0.1
        self.w1 = {k: torch.zeros((H, 3 * Dp), dtype=bf, device=dev) for k in ("actor", "critic")}
        self.w2 = {k: torch.zeros((H, H), dtype=bf, device=dev) for k in ("actor", "critic")}
        self.gpad_w = torch.zeros((self.V * self.TILE, H), dtype=torch.float32, device=dev)     # output-layer weight gradient, padded rows
        self.gpad_b = torch.zeros(self.V * self.TILE, dtype=torch.float32, device=dev)
        self.sums = torch.zeros(2, dtype=torch.float64, device=dev)                             # sum of log-ratios, loss
        self._ws = {}
        self.max_chunk = int(max_chunk)
        self.refresh()

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _buf(self, name, shape, dtype):
        t = self._ws.get(name)
        if t is None or t.shape != tuple(shape) or t.dtype != dtype:
            t = self._ws[name] = torch.empty(shape, dtype=dtype, device=self.dev)
        return t

    def _gemm(self, a, a_mn, b, b_mn, M, N, K, bias=None, act=0, mul_y=None, c32=None, accumulate=False, c16=None, row_sum=None):
        nv.check(nv.lib().vmgym_tc_gemm(a.data_ptr(), int(a_mn), a.stride(0), b.data_ptr(), int(b_mn), b.stride(0), int(M), int(N), int(K),
                                        _ptr(bias), int(act), _ptr(mul_y), mul_y.stride(0) if mul_y is not None else 0,
                                        _ptr(c32), c32.stride(0) if c32 is not None else 0, int(accumulate),
                                        _ptr(c16), c16.stride(0) if c16 is not None else 0, _ptr(row_sum), self._stream()), "vmgym_tc_gemm")

    @property
    def Dx(self):
        """Columns of the bf16 observation operand: [hi | lo | hi], each Dp wide."""
        return 3 * self.Dp

    def cast_obs(self, obs, out=None):
        """float32 observations [M, D] -> the first layer's bf16 operand [M, 3 Dp] = [hi | lo | hi]."""
        obs = obs.reshape(-1, self.D)
        if out is None:
            out = torch.empty((obs.shape[0], self.Dx), dtype=torch.bfloat16, device=self.dev)
        nv.check(nv.lib().vmgym_cast_split_bf16(obs.data_ptr(), obs.shape[0], self.D, obs.stride(0), out.data_ptr(), self.Dp, 0, self._stream()),
                 "vmgym_cast_split_bf16")
        return out

    @torch.no_grad()
    def refresh(self):
        """bf16 operand copies of the current fp32 parameters (after load_state_dict and after every optimiser step)."""
        m = self.agent.model
        lib, st = nv.lib(), self._stream()
        for k, net in (("actor", m.actor), ("critic", m.critic)):
            w1, w2 = net[0].weight, net[2].weight
            nv.check(lib.vmgym_cast_split_bf16(w1.data_ptr(), self.H, self.D, w1.stride(0), self.w1[k].data_ptr(), self.Dp, 1, st), "cast W1")
            nv.check(lib.vmgym_cast_pad_bf16(w2.data_ptr(), self.H, self.H, w2.stride(0), self.w2[k].data_ptr(), self.H, st), "cast W2")
        self.head.refresh()

    # ---- forward --------------------------------------------------------------------------------------------
    def hidden(self, which: str, x, tag: str = ""):
        """The two tanh layers of `which` ("actor" / "critic") on bf16 observations x [M, Dp] -> (a1, a2) bf16 [M, H]."""
        net = getattr(self.agent.model, which)
        M, H = x.shape[0], self.H
        a1 = self._buf(f"{which}1{tag}", (M, H), torch.bfloat16)
        a2 = self._buf(f"{which}2{tag}", (M, H), torch.bfloat16)
        self._gemm(x, 0, self.w1[which], 0, M, H, self.Dx, bias=net[0].bias, act=1, c16=a1)
        self._gemm(a1, 0, self.w2[which], 0, M, H, H, bias=net[2].bias, act=1, c16=a2)
        return a1, a2

    def values(self, obs, chunk: int | None = None):
        """critic(obs) (ppo.py:111-112) for float32 observations [M, D] -> float32 [M]."""
        obs = obs.reshape(-1, self.D)
        M = obs.shape[0]
        out = torch.empty(M, dtype=torch.float32, device=self.dev)
        chunk = chunk or self.max_chunk
        head = self.agent.model.critic[4]
        for s0 in range(0, M, chunk):
            s1 = min(M, s0 + chunk)
            x = self.cast_obs(obs[s0:s1], self._buf("xv", (s1 - s0, self.Dx), torch.bfloat16))
            _, c2 = self.hidden("critic", x, tag="v")
            nv.check(nv.lib().vmgym_value_head(c2.data_ptr(), s1 - s0, self.H, head.weight.data_ptr(), head.bias.data_ptr(),
                                               out[s0:s1].data_ptr(), self._stream()), "vmgym_value_head")
        return out

    def actor_hidden(self, obs):
        """Rollout forward: float32 observations -> the bf16 activations feeding the fused actor head."""
        x = self.cast_obs(obs, self._buf("xr", (obs.reshape(-1, self.D).shape[0], self.Dx), torch.bfloat16))
        return self.hidden("actor", x, tag="r")[1]

    # ---- one chunk of a minibatch: forward, loss, backward ------------------------------------------------------
    def begin_minibatch(self):
        self.gpad_w.zero_()
        self.gpad_b.zero_()
        self.sums.zero_()

    def forward_backward(self, x, mask, action, old_logprob, adv, old_value, ret, n_total: int):
        """x bf16 [M, 3 Dp] (cast_obs); mask int32 [M, V, 4] or None; action uint8 [M, V]; the rest float32 [M].  Accumulates the gradients
        of (sum of the samples' losses) / n_total into the flat gradient buffer / the padded output-layer scratch, and the
        sums of log-ratios / losses into self.sums."""
        ag, cfg, lib, st = self.agent, self.agent.config, nv.lib(), self._stream()
        m = ag.model
        M, H, V, A, T = x.shape[0], self.H, self.V, self.A, self.TILE
        f32, bf = torch.float32, torch.bfloat16
        a1, a2 = self.hidden("actor", x)
        c1, c2 = self.hidden("critic", x)
        lp, ent = self._buf("lp", (M, V), f32), self._buf("ent", (M, V), f32)
        use_mask = mask if cfg.masked else None
        st_m, st_s = self._buf("st_m", (M, V), f32), self._buf("st_s", (M, V), f32)
        nv.check(lib.vmgym_policy_fused_eval(a2.data_ptr(), self.head.w_pad.data_ptr(), self.head.b_pad.data_ptr(), _ptr(use_mask),
                                             action.data_ptr(), M, V, A, H, lp.data_ptr(), ent.data_ptr(), st_m.data_ptr(), st_s.data_ptr(), st),
                 "vmgym_policy_fused_eval")
        new_lp, ent_sum = lp.sum(1), ent.sum(1)                                       # ppo.py:124-126
        value = self._buf("value", (M,), f32)
        vh = m.critic[4]
        nv.check(lib.vmgym_value_head(c2.data_ptr(), M, H, vh.weight.data_ptr(), vh.bias.data_ptr(), value.data_ptr(), st), "vmgym_value_head")
        c_lp, c_v = self._buf("c_lp", (M,), f32), self._buf("c_v", (M,), f32)
        inv_n = 1.0 / float(n_total)
        nv.check(lib.vmgym_ppo_loss(new_lp.data_ptr(), old_logprob.data_ptr(), adv.data_ptr(), ent_sum.data_ptr(), value.data_ptr(),
                                    old_value.data_ptr(), ret.data_ptr(), M, float(cfg.eps_clip), float(cfg.ent_coef), float(cfg.vf_coef),
                                    int(bool(cfg.vf_loss_clip)), inv_n, c_lp.data_ptr(), c_v.data_ptr(), self.sums.data_ptr(), st), "vmgym_ppo_loss")
        # ---- actor backward ----
        g = self._buf("g", (M, V * T), bf)
        nv.check(lib.vmgym_policy_fused_grad(a2.data_ptr(), self.head.w_pad.data_ptr(), self.head.b_pad.data_ptr(), _ptr(use_mask),
                                             action.data_ptr(), M, V, A, H, c_lp.data_ptr(), -float(cfg.ent_coef) * inv_n, ent.data_ptr(),
                                             st_m.data_ptr(), st_s.data_ptr(), g.data_ptr(), g.stride(0), st), "vmgym_policy_fused_grad")
        self._gemm(g, 1, a2, 1, V * T, H, M, c32=self.gpad_w, accumulate=True, row_sum=self.gpad_b)          # dW3, db3 (padded rows)
        dz2 = self._buf("dz2", (M, H), bf)
        self._gemm(g, 0, self.head.w_pad, 1, M, H, V * T, mul_y=a2, c16=dz2)                                  # (g W3) * (1 - a2^2)
        self._layer_backward(m.actor, "actor", dz2, a1, x, M)
        # ---- critic backward ----
        dzc2 = self._buf("dzc2", (M, H), bf)
        nv.check(lib.vmgym_value_head_backward(c2.data_ptr(), M, H, vh.weight.data_ptr(), c_v.data_ptr(), dzc2.data_ptr(),
                                               vh.weight.grad.data_ptr(), vh.bias.grad.data_ptr(), st), "vmgym_value_head_backward")
        self._layer_backward(m.critic, "critic", dzc2, c1, x, M)

    def _layer_backward(self, net, which, dz2, h1, x, M):
        """dz2 = dL/d(pre-activation of layer 2) bf16 [M, H]: gradients of layers 2 and 1 of `net` (accumulated)."""
        H = self.H
        self._gemm(dz2, 1, h1, 1, H, H, M, c32=net[2].weight.grad, accumulate=True, row_sum=net[2].bias.grad)   # dW2 += dz2^T h1, db2
        dz1 = self._buf(f"dz1{which}", (M, H), torch.bfloat16)
        self._gemm(dz2, 0, self.w2[which], 1, M, H, H, mul_y=h1, c16=dz1)                                       # (dz2 W2) * (1 - h1^2)
        # dW1 += dz1^T x, db1: the hi part of x (its first D columns) is ample for a gradient
        self._gemm(dz1, 1, x, 1, H, self.D, M, c32=net[0].weight.grad, accumulate=True, row_sum=net[0].bias.grad)

    def end_minibatch(self):
        """Fold the padded output-layer gradient back into the [V * A, H] / [V * A] gradient views."""
        V, A, T, H = self.V, self.A, self.TILE, self.H
        out = self.agent.model.actor[4]
        out.weight.grad.view(V, A, H).add_(self.gpad_w.view(V, T, H)[:, :A])
        out.bias.grad.view(V, A).add_(self.gpad_b.view(V, T)[:, :A])
        return self.sums[0], self.sums[1]
