"""-m gpu: PPO-side kernels against plain torch restatements of the reference formulas (floating point: tolerances
stated per check).  Reference: src/agents/ppo.py:115-126 (masked heads), :153-155 (gating), :237-243 (GAE)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(shape="s10", n_envs=6, steps=150):
    import torch
    from vmgym import Config, VecVmEnv
    kw = dict(pms=10, vms=30, arrival_rate=0.4, service_length=30, training_steps=400, eval_steps=1000,
              reward_function="wr", allow_null_action=True) if shape == "s10" else \
        dict(pms=100, vms=300, arrival_rate=1.8182, service_length=1000, training_steps=10000, eval_steps=100000,
             reward_function="wr", allow_null_action=True)
    vec = VecVmEnv(Config(**kw), n_envs, rng="philox")
    vec.agent_step("firstfit", n_steps=steps)
    # suspend a few VMs so that running / waiting / empty rows all occur
    act = vec.vm_placement.clone()
    run = act < kw["pms"]
    act[run & (torch.rand_like(act, dtype=torch.float32) < 0.2)] = kw["pms"]
    vec.step(act)
    return vec, kw


def _ref_heads(logits, mask, action, A):
    import torch
    n, VA = logits.shape
    V = VA // A
    z = logits.double().reshape(n, V, A).masked_fill(mask, -1e7)
    logp = torch.log_softmax(z, dim=-1)
    lp = logp.gather(-1, action.long().unsqueeze(-1)).squeeze(-1).sum(1)
    ent = -(logp.exp() * logp).sum(-1).sum(1)
    return lp, ent


@pytest.mark.parametrize("shape,n_envs", [("s10", 6), ("s100", 4)])
def test_heads_forward_backward_match_torch(shape, n_envs):
    import torch
    from vmgym.ppo import PPOAgent, PPOConfig, _MaskedHeads
    vec, kw = _mk(shape, n_envs)
    agent = PPOAgent(vec, PPOConfig(hidden_size=64))
    V, A = vec.V, vec.action_dim
    torch.manual_seed(1)
    logits = (torch.randn(n_envs, V * A, device=vec.device) * 3).requires_grad_(True)
    mask = vec.get_invalid_action_mask(True)
    # sample with the on-the-fly mask; packed mask must equal the env's mask; sampled actions must be valid
    action, lp, ent, bits = agent._heads(logits.detach().contiguous(), -1.0, want_mask=True)
    W = (A + 31) // 32
    unpacked = ((bits.view(n_envs, V, W, 1) >> torch.arange(32, device=vec.device, dtype=torch.int32)) & 1).bool()
    unpacked = unpacked.reshape(n_envs, V, W * 32)[:, :, :A]
    assert torch.equal(unpacked, mask)
    assert not mask.gather(-1, action.long().unsqueeze(-1)).any(), "sampled an invalid action"
    rlp, rent = _ref_heads(logits.detach(), mask, action, A)
    assert torch.allclose(lp.double(), rlp, rtol=1e-5, atol=1e-4) and torch.allclose(ent.double(), rent, rtol=1e-5, atol=1e-4)
    # evaluate + backward through the autograd function vs torch autograd of the reference formula (fp64)
    nlp, nent = _MaskedHeads.apply(logits, bits, action, vec._ccfg(), True)
    g1, g2 = torch.randn(n_envs, device=vec.device), torch.randn(n_envs, device=vec.device)
    (nlp * g1 + nent * g2).sum().backward()
    ref_logits = logits.detach().double().requires_grad_(True)
    rlp2, rent2 = _ref_heads(ref_logits, mask, action, A)
    (rlp2 * g1.double() + rent2 * g2.double()).sum().backward()
    assert torch.allclose(nlp.double(), rlp2, rtol=1e-5, atol=1e-4)
    assert torch.allclose(logits.grad.double(), ref_logits.grad, rtol=1e-4, atol=1e-5)
    # unmasked mode (get_invalid_action_mask(False), ppo.py:152 with masked=False)
    agent.config.masked = False
    a2, lp2, ent2, _ = agent._heads(logits.detach().contiguous(), -1.0, want_mask=False)
    rlp3, rent3 = _ref_heads(logits.detach(), torch.zeros_like(mask), a2, A)
    assert torch.allclose(lp2.double(), rlp3, rtol=1e-5, atol=1e-4) and torch.allclose(ent2.double(), rent3, rtol=1e-5, atol=1e-4)


def test_gating_semantics():
    """ppo.py:153-155: rows with > 1 invalid column whose WAIT column is valid lose it with prob 1 - migration_ratio."""
    import torch
    from vmgym.ppo import PPOAgent, PPOConfig
    vec, kw = _mk("s10", 8)
    agent = PPOAgent(vec, PPOConfig(hidden_size=32))
    P, V, A = vec.P, vec.V, vec.action_dim
    logits = torch.zeros(8, V * A, device=vec.device)
    mask = vec.get_invalid_action_mask(True)
    eligible = (mask.sum(-1) > 1) & ~mask[:, :, P]

    def unpack(bits):
        W = (A + 31) // 32
        u = ((bits.view(8, V, W, 1) >> torch.arange(32, device=vec.device, dtype=torch.int32)) & 1).bool()
        return u.reshape(8, V, W * 32)[:, :, :A]
    _, _, _, b1 = agent._heads(logits, 1.0, want_mask=True)      # rand() > 1 never: mask unchanged
    assert torch.equal(unpack(b1), mask)
    _, _, _, b0 = agent._heads(logits, 0.0, want_mask=True)      # rand() > 0 (almost surely): WAIT masked for eligible rows
    m0 = unpack(b0)
    want = mask.clone()
    want[:, :, P] |= eligible
    assert torch.equal(m0, want)
    frac = []
    for _ in range(40):
        _, _, _, b = agent._heads(logits, 0.3, want_mask=True)
        frac.append((unpack(b)[:, :, P] & eligible).sum().item() / max(1, eligible.sum().item()))
    assert abs(np.mean(frac) - 0.7) < 0.05


@pytest.mark.parametrize("shape,n_envs", [("s10", 9), ("s100", 7), ("p37", 5), ("p200", 3)])
def test_mask_bits_kernel_equals_env_mask_and_heads_gating(shape, n_envs):
    """The thread-per-row mask kernel (capacity codes + byte compares) == env.py:45-53 mask, and with gating == the
    warp-per-row heads kernel on the same Philox call counter."""
    import torch
    from vmgym import Config, VecVmEnv
    from vmgym.ppo import PPOAgent, PPOConfig
    if shape in ("s10", "s100"):
        vec, kw = _mk(shape, n_envs)
    else:
        pms, vms = (37, 70) if shape == "p37" else (200, 90)
        kw = dict(pms=pms, vms=vms, arrival_rate=0.9, service_length=60, training_steps=500, eval_steps=1000,
                  reward_function="wr", allow_null_action=True)
        vec = VecVmEnv(Config(**kw), n_envs, rng="philox")
        vec.agent_step("bestfit", n_steps=130)
        act = vec.vm_placement.clone()
        run = act < pms
        act[run & (torch.rand_like(act, dtype=torch.float32) < 0.3)] = pms
        vec.step(act)
    agent = PPOAgent(vec, PPOConfig(hidden_size=32))
    P, V, A = vec.P, vec.V, vec.action_dim
    W = (A + 31) // 32

    def unpack(bits):
        u = ((bits.view(n_envs, V, W, 1) >> torch.arange(32, device=vec.device, dtype=torch.int32)) & 1).bool()
        return u.reshape(n_envs, V, W * 32)
    mask = vec.get_invalid_action_mask(True)
    b = unpack(agent._mask_bits(-1.0))
    assert torch.equal(b[:, :, :A], mask) and not b[:, :, A:].any()
    assert mask[:, :, :P].logical_not().any() and mask[:, :, :P].any()
    logits = torch.zeros(n_envs, V * A, device=vec.device)
    for ratio in (0.0, 0.3, 1.0):
        calls = agent._calls
        bm = agent._mask_bits(ratio)
        agent._calls = calls                                   # same Philox call counter for the other kernel
        _, _, _, bh = agent._heads(logits, ratio, want_mask=True)
        assert torch.equal(bm, bh), ratio
    agent.config.masked = False
    assert not agent._mask_bits(-1.0).any()


def test_sampling_distribution():
    import torch
    from vmgym import Config, VecVmEnv
    from vmgym.ppo import PPOAgent, PPOConfig
    kw = dict(pms=3, vms=5, arrival_rate=0.01, service_length=5, training_steps=100, eval_steps=100, allow_null_action=True)
    vec = VecVmEnv(Config(**kw), 512, rng="philox")
    agent = PPOAgent(vec, PPOConfig(hidden_size=8, masked=False))
    V, A = 5, 5
    row = torch.tensor([0.0, 1.0, -1.0, 2.0, 0.5], device=vec.device)
    logits = row.repeat(512, V).contiguous()
    counts = torch.zeros(A, device=vec.device)
    for _ in range(40):
        a, _, _, _ = agent._heads(logits, -1.0, want_mask=False)
        counts += torch.bincount(a.flatten().long(), minlength=A).float()
    p = torch.softmax(row, 0)
    n = counts.sum()
    chi2 = (((counts - n * p) ** 2) / (n * p)).sum().item()
    assert chi2 < 30.0, (counts / n, p)          # 4 dof: P(chi2 > 30) ~ 5e-6


@pytest.mark.parametrize("T,N", [(100, 37), (7, 5), (33, 64), (1, 3)])
def test_gae_matches_reference_loop(T, N):
    import torch
    from vmgym.ppo import gae
    g = torch.Generator().manual_seed(T * 131 + N)
    r, v, nv_ = torch.randn(T, N, generator=g), torch.randn(T, N, generator=g), torch.randn(T, N, generator=g)
    d = (torch.rand(T, N, generator=g) < 0.1)
    adv, ret = gae(r.cuda(), v.cuda(), nv_.cuda(), d.cuda(), 0.99, 0.98)
    # ppo.py:237-243, per env in float64
    ref = torch.zeros(T, N, dtype=torch.float64)
    nd = 1.0 - d.double()
    deltas = r.double() + nd * 0.99 * nv_.double() - v.double()
    run = torch.zeros(N, dtype=torch.float64)
    for i in reversed(range(T)):
        run = deltas[i] + nd[i] * 0.99 * 0.98 * run
        ref[i] = run
    assert torch.allclose(adv.cpu().double(), ref, rtol=1e-4, atol=1e-4)
    assert torch.allclose(ret.cpu().double(), ref + v.double(), rtol=1e-4, atol=1e-4)


def test_ppo_learn_smoke_and_checkpoint(tmp_path):
    import torch
    from vmgym import Config, VecVmEnv
    from vmgym.ppo import PPOAgent, PPOConfig
    kw = dict(pms=10, vms=30, arrival_rate=0.3, service_length=20, training_steps=60, eval_steps=200, reward_function="wr",
              allow_null_action=True)
    vec = VecVmEnv(Config(**kw), 16, rng="philox")
    agent = PPOAgent(vec, PPOConfig(hidden_size=64, batch_size=20, minibatch_size=5, episodes=1, env_chunk=8, lr=1e-3, kl_max=1e9))
    assert agent.mask_words == 1          # action_dim 12: the fused tcgen05 head needs 96 < action_dim <= 128, not used here
    before = [p.detach().clone() for p in agent.model.parameters()]
    agent.learn(episodes=1)
    after = list(agent.model.parameters())
    assert all(torch.isfinite(p).all() for p in after)
    assert any(not torch.equal(a, b) for a, b in zip(after, before))
    path = str(tmp_path / "ppo.pt")
    agent.save_model(path)
    assert all(k.startswith("_orig_mod.") for k in torch.load(path))
    other = PPOAgent(vec, PPOConfig(hidden_size=64))
    other.load_model(path)
    assert all(torch.equal(a, b) for a, b in zip(other.model.parameters(), agent.model.parameters()))
    vec.eval(True)
    obs, _ = vec.reset(seed=3)
    a = other.act(obs)
    assert a.shape == (16, 30)
    obs, r, term, _, info = vec.step(a)
    assert info["valid"].float().mean().item() > 0.9           # masked sampling proposes (almost) only valid actions


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 30600, 512), (200, 300, 128), (4096, 1000, 512), (1, 7, 8)])
def test_tcgen05_linear_matches_torch(M, N, K):
    """vmgym_linear_bf16 (TMA + tcgen05.mma + TMEM epilogue) vs torch on the same bf16 operands with fp32 accumulation.
    Floating point: only the summation order differs -> |diff| <= 2e-3 * (|ref| + 1) is generous for K <= 512."""
    import torch
    from vmgym.ppo import linear_bf16
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    out = linear_bf16(a, w, b)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T + b
    err = (out - ref).abs()
    assert torch.isfinite(out).all()
    assert (err <= 2e-3 * (ref.abs() + 1)).all(), float(err.max())
    out2 = linear_bf16(a, w, None)
    assert torch.allclose(out2, ref - b, rtol=2e-3, atol=2e-3)


def test_fused_actor_head_equals_gemm_plus_heads():
    """vmgym_policy_fused (logits only ever in TMEM) == vmgym_linear_bf16 -> vmgym_policy_heads on the same bf16 operands,
    mask bits and Philox stream: identical sampled actions; log-prob / entropy to fp32 summation-order tolerance."""
    import ctypes as C
    import torch
    from vmgym import _native as nv
    from vmgym.ppo import FusedActorHead, PPOAgent, PPOConfig, linear_bf16
    vec, kw = _mk("s100", 130)                       # 130 envs: a full 128-row tile and a ragged one
    agent = PPOAgent(vec, PPOConfig(hidden_size=512))
    torch.manual_seed(5)
    with torch.no_grad():
        agent.model.actor[4].weight.mul_(40.0)        # spread the logits so the softmax is not flat
        agent.model.actor[4].bias.normal_(0, 0.5)
    V, A = vec.V, vec.action_dim
    obs = vec.observe().clone()
    hidden = agent.model.actor[:4](obs).detach()
    bits = agent._mask_bits(-1.0)
    head = FusedActorHead(agent.model.actor[4], V, A)
    seed, counter = 77, 5
    act_f, lp_f, ent_f = head(hidden, bits, seed, counter)
    # unfused path on the same operands
    w3 = agent.model.actor[4].weight.detach().to(torch.bfloat16).contiguous()
    logits = linear_bf16(hidden, w3, agent.model.actor[4].bias.detach())
    n = logits.shape[0]
    act_u = torch.empty((n, V), dtype=torch.uint8, device=vec.device)
    lp_u = torch.empty(n, device=vec.device); ent_u = torch.empty(n, device=vec.device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nv.check(nv.lib().vmgym_policy_heads(C.byref(vec._ccfg()), None, bits.data_ptr(), 1, logits.data_ptr(), n, None, nv.U8, -1.0,
                                         seed, counter, act_u.data_ptr(), lp_u.data_ptr(), ent_u.data_ptr(), None, stream), "heads")
    torch.cuda.synchronize()
    assert torch.equal(act_f, act_u)
    assert torch.allclose(lp_f, lp_u, rtol=1e-4, atol=5e-3) and torch.allclose(ent_f, ent_u, rtol=1e-4, atol=5e-3)
    mask = vec.get_invalid_action_mask(True)
    assert not mask.gather(-1, act_f.long().unsqueeze(-1)).any()
    # evaluate mode: log-prob of given actions
    _, lp_e, ent_e = head(hidden, bits, seed, counter, action_in=act_f)
    assert torch.allclose(lp_e, lp_f, rtol=1e-5, atol=1e-4) and torch.allclose(ent_e, ent_f, rtol=1e-5, atol=1e-4)
    # agent-level entry point
    a2, lp2, ent2, _ = agent.fused_sample(obs)
    assert a2.shape == (130, V) and torch.isfinite(lp2).all() and not mask.gather(-1, a2.long().unsqueeze(-1)).any()


def test_ppo_learn_with_fused_rollout_s100():
    """One short training iteration at the config/100.yml shape with rollouts through the fused tcgen05 actor head."""
    import torch
    from vmgym import Config, VecVmEnv
    from vmgym.ppo import PPOAgent, PPOConfig
    kw = dict(pms=100, vms=300, arrival_rate=1.8182, service_length=1000, training_steps=12, eval_steps=100, reward_function="wr",
              allow_null_action=True)
    vec = VecVmEnv(Config(**kw), 64, rng="philox")
    agent = PPOAgent(vec, PPOConfig(hidden_size=128, batch_size=6, minibatch_size=3, episodes=1, env_chunk=32, lr=1e-4, kl_max=1e9,
                                    fused_rollout=True))
    before = [p.detach().clone() for p in agent.model.parameters()]
    agent.learn(episodes=1)
    assert all(torch.isfinite(p).all() for p in agent.model.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(agent.model.parameters(), before))
    assert vec.counters()["place_actions"].sum() > 0


def test_ppo_eval_rollout_fused_and_unfused():
    """PPOAgent.rollout (eval loop with act): both forward paths run, only valid actions reach the env, rewards accumulate."""
    import torch
    from vmgym import Config, VecVmEnv
    from vmgym.ppo import PPOAgent, PPOConfig
    kw = dict(pms=100, vms=300, arrival_rate=1.8182, service_length=1000, training_steps=10000, eval_steps=100000,
              reward_function="wr", allow_null_action=True)
    for fused in (True, False):
        vec = VecVmEnv(Config(**kw), 32, rng="philox")
        vec.eval(True)
        agent = PPOAgent(vec, PPOConfig(hidden_size=128, migration_ratio=0.002))
        ret = agent.rollout(12, fused=fused)
        c = vec.counters()
        assert ret.shape == (32,) and torch.isfinite(ret).all() and (ret <= 0).all()
        assert c["timestep"].min() == 13 and c["place_actions"].sum() > 0
        assert torch.allclose(ret, torch.from_numpy(c["episode_return"]).to(ret.device))


@pytest.mark.parametrize("which", ["ppo", "drlvmp"])
def test_learned_agents_test_returns_a_record(which, tmp_path):
    """Base.test (base.py:63-124) for the learned agents: act/step loop until the episode ends, Record summary from the
    device-side sums (steps, requests, per-VM keys), JSON output."""
    import json
    import torch
    from vmgym import Config, VmEnv
    kw = dict(pms=10, vms=30, arrival_rate=0.4, service_length=20, training_steps=50, eval_steps=80, reward_function="wr",
              allow_null_action=True, seed=3)
    env = VmEnv(Config(**kw), rng="philox")
    torch.manual_seed(0)
    if which == "ppo":
        from src.agents.ppo import PPOAgent, PPOConfig
        agent = PPOAgent(env, PPOConfig(hidden_size=32, migration_ratio=0.3))
    else:
        from src.agents.drlvmp import DRLVMPAgent, DRLVMPConfig
        agent = DRLVMPAgent(env, DRLVMPConfig(hidden_size=16))
    out = tmp_path / "r.json"
    rec = agent.test(output=str(out))
    s = rec.get_summary()
    assert len(s) == 22 and s["total requests"] > 0 and s["total place actions"] >= 0
    assert float(rec.raw["steps"][0]) == 80
    c = env.vec.counters()
    assert s["total served VMs"] == int(c["served_requests"][0]) and s["total suspend actions"] == int(c["suspend_actions"][0])
    assert json.load(open(out))["summary"] == s
    agent.set_log("job", None); agent.end_log()


@pytest.mark.gpu
def test_out_linear_backward_matches_torch_linear():
    """The update's output layer (bias gradient folded into the weight-gradient GEMM) == F.linear's autograd, fp32 matmuls."""
    import torch
    import torch.nn.functional as F
    from vmgym.ppo import _OutLinear
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        g = torch.Generator(device="cuda").manual_seed(5)
        h = torch.randn(257, 48, device="cuda", generator=g, requires_grad=True)
        w = torch.randn(1234, 48, device="cuda", generator=g, requires_grad=True)
        b = torch.randn(1234, device="cuda", generator=g, requires_grad=True)
        up = torch.randn(257, 1234, device="cuda", generator=g)
        out1 = _OutLinear.apply(h, w, b)
        g1 = torch.autograd.grad((out1 * up).sum(), (h, w, b))
        out2 = F.linear(h, w, b)
        g2 = torch.autograd.grad((out2 * up).sum(), (h, w, b))
        assert torch.equal(out1, out2)
        for a, c in zip(g1, g2):
            assert a.shape == c.shape and a.is_contiguous()
            assert torch.allclose(a, c, rtol=1e-4, atol=1e-4)          # summation order only
    finally:
        torch.set_float32_matmul_precision(prev)
