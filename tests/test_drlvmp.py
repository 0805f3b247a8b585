"""DRL-VMP rollout side against golden vectors recorded from the unmodified reference (tests/golden/drlvmp.npz,
made by tests/golden/make_golden_drlvmp.py): network forward on CPU (not gpu), heuristics kernel and act loop on GPU."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "drlvmp.npz")


def _load():
    z = np.load(GOLD)
    return {k: z[k] for k in z.files}


def test_network_matches_reference_q_values():
    """drlvmp.py:355-372 — same state_dict keys, same q-values (fp32, CPU): |dq| <= 1e-4 on q ~ 100."""
    from vmgym.drlvmp import Network
    g = _load()
    net = Network(110, 24, 4, 51, torch.linspace(0.0, 200.0, 51))
    sd = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("net.")}
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    with torch.no_grad():
        q = net(torch.from_numpy(g["net_x"]))
    assert torch.allclose(q, torch.from_numpy(g["net_q"]), atol=1e-4, rtol=1e-6)
    assert torch.equal(q.argmax(1), torch.from_numpy(g["net_q"]).argmax(1))


@pytest.mark.gpu
@pytest.mark.parametrize("base,P,V", [("10", 10, 30), ("100", 100, 300)])
def test_heuristics_kernel_matches_reference(base, P, V):
    """_get_worstfit/_dot/_norm2/_bestfit_action (drlvmp.py:549-617) on the reference's own observations."""
    from vmgym import Config, VecVmEnv
    from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig
    g = _load()
    vec = VecVmEnv(Config(pms=P, vms=V, allow_null_action=True, training_steps=10, eval_steps=10), 1, rng="philox")
    agent = DRLVMPAgent(vec, DRLVMPConfig(hidden_size=8))
    obs = torch.from_numpy(g[f"h{base}_obs"]).cuda()
    pm = agent.heuristic(obs.contiguous(), torch.from_numpy(g[f"h{base}_v"]).cuda(), torch.from_numpy(g[f"h{base}_choice"]).cuda())
    got, want, ch = pm.cpu().numpy(), g[f"h{base}_pm"], g[f"h{base}_choice"]
    o = g[f"h{base}_obs"]
    key = o[:, 3 * V:3 * V + P] + o[:, 3 * V + P:]                     # cpu + memory, float32 (drlvmp.py:556,575)
    n_tie = 0
    for i in np.flatnonzero(got != want):
        # torch.argsort's order among EQUAL keys is unspecified (SURVEY §8c "DRL-VMP"): a different PM is accepted only
        # for the two sort-based heuristics, only when both PMs carry the same key, and ours must follow the pinned rule
        assert ch[i] in (0, 3) and got[i] >= 0 and want[i] >= 0, (int(i), int(ch[i]), int(got[i]), int(want[i]))
        assert key[i, got[i]] == key[i, want[i]], (int(i), int(ch[i]), int(got[i]), int(want[i]))
        same = np.flatnonzero(key[i] == key[i, got[i]])
        v = g[f"h{base}_v"][i]
        fits = [p for p in same if np.float32(o[i, 3 * V + p] + o[i, V + v]) <= 1 and np.float32(o[i, 3 * V + P + p] + o[i, 2 * V + v]) <= 1]
        assert got[i] == (min(fits) if ch[i] == 0 else max(fits))      # worst-fit: lowest index, best-fit: highest
        n_tie += 1
    assert n_tie < 0.5 * len(got)


@pytest.mark.gpu
def test_act_loop_equals_sequential_restatement():
    """act (drlvmp.py:504-512): per waiting VM in slot order, network choice on the working observation, heuristic on
    it, placement written back — batched over envs == one env at a time."""
    from vmgym import Config, VecVmEnv
    from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig
    torch.manual_seed(0)
    vec = VecVmEnv(Config(pms=10, vms=30, arrival_rate=0.5, service_length=40, allow_null_action=True, training_steps=500,
                          eval_steps=500), 7, rng="philox")
    vec.agent_step("firstfit", n_steps=60)
    act = vec.vm_placement.clone()
    act[(act < 10) & (torch.rand_like(act, dtype=torch.float32) < 0.5)] = 10        # suspend some -> waiting VMs that fit
    obs, *_ = vec.step(act)
    obs = obs.clone()
    agent = DRLVMPAgent(vec, DRLVMPConfig(hidden_size=16))
    batched = agent.act(obs).cpu().numpy()                       # incremental feature update (default)
    assert np.array_equal(batched, agent.act(obs, incremental=False).cpu().numpy())
    assert np.array_equal(batched, agent.act(obs, refresh=3).cpu().numpy())
    assert np.array_equal(batched, agent.act(obs, graph=True).cpu().numpy())            # CUDA-graph replay per waiting VM
    assert np.array_equal(batched, agent.act(obs, graph=True, refresh=2).cpu().numpy())  # cached graph, second call
    assert np.array_equal(batched, agent.act(obs, graph=True, fused=False).cpu().numpy())  # graph of torch ops + choice kernel
    assert np.array_equal(batched, agent.act(obs, graph=True, fused=True, refresh=5).cpu().numpy())  # 2 GEMMs + vmgym_drlvmp_iter
    for i in range(7):
        o = obs[i:i + 1].clone()
        for v in torch.nonzero(obs[i, :30] == 10.0).flatten().tolist():
            choice = agent.dqn(o).argmax(1).to(torch.int32)
            pm = agent.heuristic(o.contiguous(), torch.tensor([v], dtype=torch.int32, device="cuda"), choice)
            if pm.item() >= 0:
                o[0, v] = float(pm.item())
        assert np.array_equal(batched[i], o[0, :30].long().cpu().numpy()), i
    one = agent.act(obs[3].cpu().numpy())
    assert one.dtype == np.int64 and np.array_equal(one, batched[3])


@pytest.mark.gpu
def test_device_segment_trees_match_reference():
    """SumSegmentTree / MinSegmentTree (src/segment_tree.py): batched stores, sum(), min(), retrieve() — bit-identical fp64."""
    from vmgym.drlvmp import DeviceSegmentTrees
    g = _load()
    trees = DeviceSegmentTrees(int(g["st_cap"]))
    for b in range(g["st_idx"].shape[0]):
        trees.set(g["st_idx"][b], g["st_val"][b])
        assert trees.sum().item() == g["st_sum"][b]
        assert trees.min().item() == g["st_min"][b]
        got = trees.retrieve(g["st_ub"][b]).cpu().numpy()
        assert np.array_equal(got, g["st_ret"][b]), b
    with pytest.raises(AssertionError):
        DeviceSegmentTrees(1000)


# ---- training internals (vmgym/drlvmp_train.py) against golden vectors from the reference classes ---------------------
def _train_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "drlvmp_train.npz"))


@pytest.mark.gpu
def test_c51_loss_matches_reference():
    """_compute_dqn_loss (drlvmp.py:661-706) for the 1-step and the n-step gamma: projection kernel + torch layers vs the
    reference agent's own losses (float32: 2e-5 relative, 1e-6 absolute — the projection itself is checked exactly below)."""
    import torch
    from vmgym.drlvmp import Network
    from vmgym.drlvmp_train import c51_project, dqn_loss
    z = _train_golden()
    support = torch.linspace(0.0, 200.0, 51, device="cuda")
    nets = {}
    for name in ("dqn", "tgt"):
        net = Network(110, 24, 4, 51, support).cuda()
        net.load_state_dict({k[len(f"c51_{name}."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"c51_{name}.")})
        nets[name] = net
    samples = {k: torch.from_numpy(z["c51_" + k]).cuda() for k in ("obs", "next_obs", "acts", "rews", "done")}
    for tag in ("g1", "g3"):
        loss = dqn_loss(nets["dqn"], nets["tgt"], samples, float(z["c51_gamma_" + tag]), 0.0, 200.0)
        assert torch.allclose(loss.cpu(), torch.from_numpy(z["c51_loss_" + tag]), rtol=2e-5, atol=1e-6), tag
    assert float(loss[3].detach()) == 0.0   # reward 250: every atom clamps onto the top grid point and the mass is dropped (:683-699)
    # the projection alone, bit for bit against the reference's formula evaluated by torch on the CPU
    g = torch.Generator().manual_seed(9)
    B, A = 257, 51
    nd = torch.softmax(torch.randn(B, A, generator=g), -1).clamp(min=1e-3)
    rew = (torch.rand(B, generator=g) * 260 - 20).float()
    rew[:40] = torch.arange(40).float() * 4.0                                # integral b values
    done = (torch.rand(B, generator=g) < 0.3).int()
    sup = torch.linspace(0.0, 200.0, A)
    for gamma in (0.99, 0.99 ** 3):
        t_z = (rew.reshape(-1, 1) + (1 - done.reshape(-1, 1)) * gamma * sup).clamp(min=0.0, max=200.0)
        b = (t_z - 0.0) / (200.0 / (A - 1))
        lo, up = b.floor().long(), b.ceil().long()
        off = torch.linspace(0, (B - 1) * A, B).long().unsqueeze(1).expand(B, A)
        want = torch.zeros(B, A)
        want.view(-1).index_add_(0, (lo + off).view(-1), (nd * (up.float() - b)).view(-1))
        want.view(-1).index_add_(0, (up + off).view(-1), (nd * (b - lo.float())).view(-1))
        got = c51_project(nd.cuda(), rew.cuda(), done.cuda(), sup.cuda(), gamma, 0.0, 200.0).cpu()
        assert torch.equal(got, want), gamma


@pytest.mark.gpu
def test_n_step_buffer_matches_reference():
    """ReplayBuffer.store / _get_n_step_info (drlvmp.py:46-113), one env: the ring contents after 40 transitions with dones."""
    import torch
    from vmgym.drlvmp_train import NStepReplay
    z = _train_golden()
    rb = NStepReplay(4, 32, num_envs=1, n_step=3, gamma=0.99)
    obs = torch.from_numpy(z["ns_obs"]).cuda()
    for t in range(z["ns_acts"].size):
        rb.store(obs[t:t + 1], torch.tensor([int(z["ns_acts"][t])], device="cuda"), torch.tensor([z["ns_rews"][t]], device="cuda"),
                 obs[t + 1:t + 2], torch.tensor([bool(z["ns_dones"][t])], device="cuda"))
    assert rb.ptr == int(z["ns_ptr"]) and rb.size == int(z["ns_size"])
    assert torch.equal(rb.obs_buf.cpu(), torch.from_numpy(z["ns_obs_buf"]))
    assert torch.equal(rb.next_obs_buf.cpu(), torch.from_numpy(z["ns_next_obs_buf"]))
    assert torch.equal(rb.acts_buf.cpu(), torch.from_numpy(z["ns_acts_buf"]))
    assert torch.equal(rb.rews_buf.cpu(), torch.from_numpy(z["ns_rews_buf"]))
    assert torch.equal(rb.done_buf.cpu(), torch.from_numpy(z["ns_done_buf"]))


@pytest.mark.gpu
def test_prioritized_replay_matches_reference():
    """PrioritizedReplayBuffer (drlvmp.py:118-241): stratified sampling with the reference's uniforms picks the same
    indices, importance weights agree to 1e-12, priorities / max_priority / tree sums evolve identically."""
    import torch
    from vmgym.drlvmp_train import PrioritizedReplay
    z = _train_golden()
    per = PrioritizedReplay(4, 48, num_envs=1, alpha=float(z["per_alpha"]))
    o = torch.zeros(1, 4, device="cuda")
    for t in range(int(z["per_n"])):
        per.store(o, torch.tensor([1], device="cuda"), torch.tensor([0.0], device="cuda"), o, torch.tensor([False], device="cuda"))
    for r in range(z["per_u"].shape[0]):
        s = per.sample_batch(8, beta=0.5 + 0.05 * r, u=z["per_u"][r])
        assert np.array_equal(s["indices"].cpu().numpy(), z["per_idx"][r]), r
        assert np.allclose(s["weights"].cpu().numpy(), z["per_w"][r], rtol=1e-12, atol=0), r
        per.update_priorities(s["indices"], torch.from_numpy(z["per_upd_pri"][r]))
    assert float(per.max_priority) == float(z["per_max_priority"])
    # priority ** alpha is CUDA's fp64 pow here and libm's in the reference: 2 ulp
    assert float(per.trees.sum()) == pytest.approx(float(z["per_sum"]), rel=1e-14)
    assert float(per.trees.min()) == pytest.approx(float(z["per_min"]), rel=1e-14)


@pytest.mark.gpu
def test_drlvmp_learn_runs_and_updates():
    """DRLVMPAgent.learn over a batch of envs: buffers fill, losses are finite, the online net changes, the target net is
    synchronised every target_update optimisation steps."""
    import torch
    from vmgym import Config, VecVmEnv
    from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig
    kw = dict(pms=10, vms=30, arrival_rate=0.4, service_length=30, training_steps=60, eval_steps=100, reward_function="wr",
              allow_null_action=True)
    vec = VecVmEnv(Config(**kw), 8, rng="philox")
    torch.manual_seed(0)
    agent = DRLVMPAgent(vec, DRLVMPConfig(hidden_size=32, batch_size=16, memory_size=512, episodes=2, n_step=3, target_update=5))
    before = [p.detach().clone() for p in agent.dqn.parameters()]
    returns = agent.learn(episodes=2)
    tr = agent.trainer
    assert returns.shape == (2, 8) and np.isfinite(returns).all()
    assert len(tr.memory) == len(tr.memory_n) == min(512, 8 * (2 * 60 - 2))
    assert len(tr.losses) > 50 and all(torch.isfinite(l) for l in tr.losses)
    assert any(not torch.equal(a, b) for a, b in zip(agent.dqn.parameters(), before))
    assert vec.counters()["place_actions"].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["act_s10", "act_s100"])
@pytest.mark.parametrize("mode", ["eager", "graph", "fused"])
def test_whole_act_loop_matches_reference_agent(case, mode):
    """DRLVMPAgent.act (drlvmp.py:504-530) end to end against actions recorded from the UNMODIFIED reference agent
    (tests/golden/make_golden_drlvmp_act.py: random network, observations from a reference run): the same placement vector for every
    observation, in every execution mode of the device loop.  The recorded decisions have a q-value margin >= 0.04 and none hinges on a tie of equal keys (unspecified order in torch), far
    above the 1e-4 network tolerance, so the comparison is exact."""
    import json
    from vmgym import Config, VecVmEnv
    from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig
    z = np.load(os.path.join(os.path.dirname(GOLD), "drlvmp_act.npz"))
    fx = {k[len(case) + 1:]: z[k] for k in z.files if k.startswith(case + ".")}
    cfg = json.loads(str(fx["cfg_json"]))
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        vec = VecVmEnv(Config(**cfg), 1, rng="philox")
        agent = DRLVMPAgent(vec, DRLVMPConfig(hidden_size=int(fx["hidden"])))
        agent.dqn.load_state_dict({k[3:]: torch.from_numpy(v).cuda() for k, v in fx.items() if k.startswith("sd.")}, strict=True)
        agent.eval()
        obs = torch.from_numpy(fx["obs"]).cuda()
        n = obs.shape[0]
        reps = 1 if mode == "eager" else (64 + n - 1) // n               # the graph paths want batches of >= 64 rows
        batch = obs.repeat(reps, 1)
        kw = dict(eager=dict(graph=False), graph=dict(graph=True, fused=False), fused=dict(graph=True, fused=True))[mode]
        act = agent.act(batch, **kw).cpu().numpy()
        want = np.tile(fx["action"], (reps, 1))
        assert np.array_equal(act, want), f"{(act != want).sum()} of {want.size} placements differ from the reference agent's"
        assert float(fx["margin"][fx["choices"] >= 0].min()) > 0.01
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
