"""-m gpu: parity of the CUDA env path (through the C ABI, via the Python mirror) against the oracle and the
golden fixtures recorded from the reference.

Bar (BASELINE.json north_star): placements, validity flags, queue contents (slot arrays), remaining runtimes and
episode counters bit-exact; fp64 PM accumulators bit-exact; float32 observations bit-exact; rewards within 1e-6
relative (they are in fact bit-exact for wr/ut and ~1e-14 for kl, whose log() differs from glibc's by <= 1 ulp).
"""
import numpy as np
import pytest

import golden_util as gu
import vmoracle as vo

pytestmark = pytest.mark.gpu

REWARD_RTOL = 1e-6   # the tolerance north_star states (fp64 on both sides; see module docstring)


def _torch():
    import torch
    return torch


def _cfg(**kw):
    from vmgym import Config
    return Config(**kw)


def _make_cuda_env(cfg, trace_steps, trace_adm):
    from vmgym import VmEnv
    return VmEnv(_cfg(**cfg), trace_steps=trace_steps, max_admissions=trace_adm)


@pytest.mark.parametrize("name", gu.fixture_names())
def test_cuda_env_replays_reference_fixture(name):
    fx = gu.load(name)
    tie = "numpy_introsort" if str(fx["tiebreak"]) == "numpy_introsort" else "stable"
    steps = 600 if fx["cfg"]["pms"] >= 100 else 1000
    if fx["reset_at"].size:
        steps = None          # episode fixtures: run through the resets
    gu.replay(fx, _make_cuda_env, reward_rtol=REWARD_RTOL, max_steps=steps)
    del tie


def _oracle_batch(cfg_kw, seeds, steps, adm):
    envs = []
    for s in seeds:
        kw = dict(cfg_kw, seed=int(s))
        envs.append(vo.OracleVmEnv(vo.OracleConfig(**kw), trace_steps=steps, trace_adm=adm))
    return envs


_SZ32 = (np.arange(101) / 100.0).astype(np.float32)


def _expected_caps(load64):
    """max size code k with float32(load) + float32(k/100) <= 1 in float32 arithmetic (the agents' fit test)."""
    x = load64.astype(np.float32)[:, None]
    fits = (x + _SZ32[None, :]).astype(np.float32) <= np.float32(1.0)
    return fits.sum(1).astype(np.int64) - 1


def _compare_state(vec, oracles, t, check_reward=None):
    cnt = vec.counters()
    L = vec._layout
    caps = vec.state[:, L.off_capacity:L.off_capacity + 2 * vec.P].cpu().numpy().view(np.uint16)
    slot = cnt["slot_counts"]
    for i, o in enumerate(oracles):
        s = vec.state_dict_host(i)
        so = o.state()
        for k in ("vm_placement", "vm_remaining_runtime", "vm_suspended"):
            assert np.array_equal(s[k], so[k]), f"{k} env {i} step {t}"
        assert np.array_equal(s["vm_cpu"], so["vm_cpu"]) and np.array_equal(s["vm_memory"], so["vm_memory"]), (i, t)
        assert s["cpu"].tobytes() == so["cpu"].tobytes(), f"fp64 cpu env {i} step {t}"
        assert s["memory"].tobytes() == so["memory"].tobytes(), f"fp64 memory env {i} step {t}"
        for k in ("timestep", "total_requests", "served_requests", "dropped_requests", "suspend_action", "place_action",
                  "arr_cursor", "adm_cursor", "trace_exhausted"):
            assert s[k] == so[k], f"{k} env {i} step {t}: {s[k]} vs {so[k]}"
        # derived caches kept in the record: per-PM capacity codes and the slot counters
        assert np.array_equal(caps[i] & 0xff, _expected_caps(so["cpu"])), f"cpu capacity codes env {i} step {t}"
        assert np.array_equal(caps[i] >> 8, _expected_caps(so["memory"])), f"memory capacity codes env {i} step {t}"
        assert (slot[i] & 0xffff) == int((so["vm_placement"] == vec.P).sum()), f"n_waiting env {i} step {t}"
        assert (slot[i] >> 16) == int((so["vm_placement"] == vec.P + 1).sum()), f"n_empty env {i} step {t}"
    return cnt


SHAPES = {
    "s10": dict(pms=10, vms=30, arrival_rate=0.3, service_length=30, training_steps=400, eval_steps=100000,
                reward_function="kl", allow_null_action=True),
    "s10wr": dict(pms=10, vms=30, arrival_rate=0.3, service_length=30, training_steps=400, eval_steps=100000,
                  reward_function="wr", allow_null_action=True),       # config/10.yml shape on the specialised kernels
    "s100": dict(pms=100, vms=300, arrival_rate=1.8182, service_length=1000, training_steps=10000, eval_steps=100000,
                 reward_function="wr", allow_null_action=True),
    "odd": dict(pms=33, vms=65, arrival_rate=1.0, service_length=50, training_steps=10000, eval_steps=100000,
                reward_function="ut", beta=0.3, allow_null_action=False, sequence="highuniform", cap_target_util=False),
    "s1000": dict(pms=1000, vms=3000, arrival_rate=1.6, service_length=1000, training_steps=10000, eval_steps=100000,
                  reward_function="wr", allow_null_action=True, sequence="highuniform"),      # BASELINE config 5 shape
    "p253": dict(pms=253, vms=61, arrival_rate=2.0, service_length=25, training_steps=10000, eval_steps=100000,
                 reward_function="kl", allow_null_action=True),                                # largest byte-placement layout
    "p254": dict(pms=254, vms=61, arrival_rate=2.0, service_length=25, training_steps=10000, eval_steps=100000,
                 reward_function="ut", allow_null_action=True),                                # smallest u16-placement layout
    "wide": dict(pms=300, vms=700, arrival_rate=4.0, service_length=120, training_steps=10000, eval_steps=100000,
                 reward_function="kl", allow_null_action=True, sequence="lowuniform"),
    # u16 placements with a short service time: departures, re-admissions and re-placements every step (team-mode kernel)
    "big": dict(pms=400, vms=1100, arrival_rate=9.0, service_length=40, training_steps=10000, eval_steps=100000,
                reward_function="wr", allow_null_action=True, sequence="highuniform"),
}


@pytest.mark.parametrize("shape,n_envs,steps", [("s10", 37, 450), ("s100", 8, 300), ("odd", 16, 400), ("wide", 4, 250),
                                                ("p253", 3, 200), ("p254", 3, 200), ("s10", 1, 120)])
@pytest.mark.parametrize("bulk", [7, 3, 0])      # bits: bulk loads | bulk stores | programmatic dependent launch
def test_batched_step_matches_oracle_random_actions(shape, n_envs, steps, bulk):
    """N envs with different seeds, adversarial random action streams (out-of-range values, suspend storms,
    simultaneous placements on one PM, NULL-slot actions): every env equals its own oracle after every step."""
    torch = _torch()
    from vmgym import VecVmEnv
    from vmgym import _native as nv
    kw = SHAPES[shape]
    nv.lib().vmgym_set_tuning(0, bulk)
    try:
        seeds = 100 + 7 * np.arange(n_envs)
        vec = VecVmEnv(_cfg(**kw), n_envs, seeds=seeds, trace_steps=steps + 4, max_admissions=20000)
        oracles = _oracle_batch(kw, seeds, steps + 4, 20000)
        P, V, A = kw["pms"], kw["vms"], vec.action_dim
        rng = np.random.default_rng(99)
        obs = vec.obs.cpu().numpy()
        for i, o in enumerate(oracles):
            assert np.array_equal(obs[i], o._obs())
        for t in range(steps):
            place = np.stack([o.state()["vm_placement"] for o in oracles])
            ff = np.stack([vo.firstfit_act(P, V, o._obs()) for o in oracles])
            u = rng.random((n_envs, V))
            act = np.where(u < 0.5, ff, place)
            act = np.where(u > 0.93, rng.integers(0, A + 2, size=(n_envs, V)), act)
            act = np.where((u > 0.88) & (u <= 0.93), P, act)
            dtype = [torch.int64, torch.int16, torch.uint8][t % 3] if P <= 253 else [torch.int64, torch.int16][t % 2]
            obs_d, rew_d, term_d, _, info = vec.step(torch.from_numpy(act).to(vec.device).to(dtype))
            obs_h, rew_h, term_h, valid_h = obs_d.cpu().numpy(), rew_d.cpu().numpy(), term_d.cpu().numpy(), info["valid"].cpu().numpy()
            for i, o in enumerate(oracles):
                a = act[i].astype(np.int64)
                if dtype == torch.uint8:
                    a = a.astype(np.uint8).astype(np.int64)
                o_obs, o_r, o_term, _, o_info = o.step(a)
                assert np.array_equal(valid_h[i], o_info["valid"].astype(np.uint8)), (i, t)
                assert obs_h[i].tobytes() == o_obs.tobytes(), (i, t)
                assert rew_h[i] == pytest.approx(o_r, rel=REWARD_RTOL, abs=1e-300), (i, t)
                assert bool(term_h[i]) == o_term
            if t % 25 == 0 or t == steps - 1:
                _compare_state(vec, oracles, t)
    finally:
        nv.lib().vmgym_set_tuning(0, 7)


@pytest.mark.parametrize("agent,tie", [("firstfit", "stable"), ("bestfit", "stable"), ("bestfit", "numpy_introsort")])
@pytest.mark.parametrize("shape,n_envs,steps,chunk", [("s100", 6, 1300, 1), ("s100", 6, 1300, 64), ("s10", 20, 380, 7),
                                                       ("odd", 9, 500, 25), ("s1000", 2, 150, 50), ("big", 3, 240, 40)])
@pytest.mark.parametrize("vectors", [True, False])
def test_fused_agent_step_matches_oracle_rollout(agent, tie, shape, n_envs, steps, chunk, vectors):
    """agent.act + env.step fused in one kernel (chunk steps per launch) == oracle act()/step() loop.
    vectors=False drops the per-slot action/valid outputs, which lets the kernel skip act()+apply on QUIET steps."""
    from vmgym import VecVmEnv
    kw = SHAPES[shape]
    seeds = 5 + 3 * np.arange(n_envs)
    vec = VecVmEnv(_cfg(**kw), n_envs, seeds=seeds, trace_steps=steps + 4, max_admissions=30000, tiebreak=tie)
    oracles = _oracle_batch(kw, seeds, steps + 4, 30000)
    P, V = kw["pms"], kw["vms"]
    otie = vo.TIE_NUMPY_INTROSORT if tie == "numpy_introsort" else vo.TIE_STABLE
    returns = np.zeros(n_envs)
    done = 0
    while done < steps:
        n = min(chunk, steps - done)
        vec.agent_step(agent, n_steps=n, want_stats=True, want_action=vectors, want_valid=vectors)
        act_d = vec.agent_action.cpu().numpy().astype(np.int64)
        obs_h, rew_h = vec.obs.cpu().numpy(), vec.reward.cpu().numpy()
        for i, o in enumerate(oracles):
            for k in range(n):
                ob = o._obs()
                a = vo.firstfit_act(P, V, ob) if agent == "firstfit" else vo.bestfit_act(P, V, ob, otie)
                o_obs, o_r, o_term, _, _ = o.step(a)
                returns[i] += o_r
                if o_term:
                    break
            assert not vectors or np.array_equal(act_d[i], a), (i, done)
            assert obs_h[i].tobytes() == o_obs.tobytes(), (i, done)
            assert rew_h[i] == pytest.approx(o_r, rel=REWARD_RTOL, abs=1e-300)
        done += n
        if done % 256 < chunk or done == steps:
            cnt = _compare_state(vec, oracles, done)
            assert np.allclose(cnt["episode_return"], returns, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("team_warps", [1, 3, 8])
@pytest.mark.parametrize("shape,agent", [("big", "bestfit"), ("wide", "firstfit"), ("p254", "bestfit")])
def test_team_mode_widths_match_oracle(team_warps, shape, agent):
    """Large shapes (u16 placements) run one env per CTA with helper warps for the bulk phases: every team width gives the
    oracle's trajectory (per-slot state, observation bytes, rewards), single- and multi-step launches."""
    from vmgym import VecVmEnv
    from vmgym import _native as nv
    kw = SHAPES[shape]
    n_envs, steps = 5, 230
    seeds = 11 + 5 * np.arange(n_envs)
    nv.lib().vmgym_set_tuning(team_warps, 7)
    try:
        vec = VecVmEnv(_cfg(**kw), n_envs, seeds=seeds, trace_steps=steps + 4, max_admissions=30000)
        oracles = _oracle_batch(kw, seeds, steps + 4, 30000)
        P, V = kw["pms"], kw["vms"]
        done = 0
        for n in [1, 1, 1, 7, 50, 1, 90, 79]:
            vec.agent_step(agent, n_steps=n, want_action=False, want_valid=False)
            obs_h, rew_h = vec.obs.cpu().numpy(), vec.reward.cpu().numpy()
            for i, o in enumerate(oracles):
                for _ in range(n):
                    ob = o._obs()
                    a = vo.firstfit_act(P, V, ob) if agent == "firstfit" else vo.bestfit_act(P, V, ob, vo.TIE_STABLE)
                    o_obs, o_r, _, _, _ = o.step(a)
                assert obs_h[i].tobytes() == o_obs.tobytes(), (i, done)
                assert rew_h[i] == pytest.approx(o_r, rel=REWARD_RTOL, abs=1e-300)
            done += n
            _compare_state(vec, oracles, done)
        assert done == steps
        c = vec.counters()
        assert int(c["served_requests"].sum()) > 0 and int(c["place_actions"].sum()) > 0
    finally:
        nv.lib().vmgym_set_tuning(0, 7)


@pytest.mark.parametrize("shape,n_envs,agent,rng", [("s100", 9000, "bestfit", "philox"), ("s10wr", 12000, "firstfit", "philox"),
                                                     ("odd", 7000, "bestfit", "numpy"), ("s100", 7000, "external", "philox")])
def test_double_buffered_multi_round_launch(shape, n_envs, agent, rng):
    """More envs than resident warps: the launch walks several envs per warp with double-buffered records (next record
    prefetched, previous write-back draining).  Same records, observations and rewards as the single-buffered kernel
    (use_bulk bit 4), and sampled envs equal the oracle."""
    torch = _torch()
    from vmgym import VecVmEnv
    from vmgym import _native as nv
    kw = SHAPES[shape]
    steps = 45
    seeds = 1000 + np.arange(n_envs)
    sample = [0, 1, n_envs // 3, n_envs // 2 + 1, n_envs - 2, n_envs - 1]
    results = []
    try:
        for bits in (7 + 32, 7 + 16):              # double-buffered for any record size / never
            nv.lib().vmgym_set_tuning(0, bits)
            if rng == "philox":
                vec = VecVmEnv(_cfg(**kw), n_envs, rng="philox", seeds=seeds)
            else:
                vec = VecVmEnv(_cfg(**kw), n_envs, seeds=seeds, trace_steps=steps + 4, max_admissions=4000)
            rewards = []
            for n in (1, 3, 1, 17, 1, 22):
                if agent == "external":
                    for _ in range(n):
                        a = vec.vm_placement.clone()
                        a[:, ::3] = vec.P                                     # suspend / keep waiting every third slot
                        a[:, 1::7] = (torch.arange(a[:, 1::7].shape[1], device=a.device) % vec.P).to(a.dtype)
                        vec.step(a)
                        rewards.append(vec.reward.clone())
                else:
                    vec.agent_step(agent, n_steps=n, want_action=False, want_valid=False)
                    rewards.append(vec.reward.clone())
            results.append((vec.state.clone(), vec.obs.clone(), torch.stack(rewards), vec))
    finally:
        nv.lib().vmgym_set_tuning(0, 7)
    (s_a, o_a, r_a, vec_a), (s_b, o_b, r_b, _) = results
    assert torch.equal(s_a, s_b) and torch.equal(o_a, o_b) and torch.equal(r_a, r_b)
    if agent != "external":
        # sampled envs against the oracle (an env's trajectory does not depend on the batch it is in)
        P, V = kw["pms"], kw["vms"]
        obs_h = o_a.cpu().numpy()
        for i in sample:
            if rng == "philox":
                ka, ta, ks, ts, lo, hi = vec_a.philox_tables
                o = vo.OracleVmEnv(vo.OracleConfig(**dict(kw, seed=0)), trace_steps=4, trace_adm=4)
                o.reset(trace=vo.philox_trace(int(seeds[i]), steps + 4, 4000, ka, ta, ks, ts, lo, hi))
            else:
                o = _oracle_batch(kw, seeds[i:i + 1], steps + 4, 4000)[0]
            o.rollout(vo.AGENT_FIRSTFIT if agent == "firstfit" else vo.AGENT_BESTFIT, steps)
            assert obs_h[i].tobytes() == o._obs().tobytes(), i


@pytest.mark.parametrize("name", ["s100_firstfit_wr", "s100_bestfit_stable_wr", "s100_bestfit_introsort_wr",
                                  "s10_bestfit_stable_ut", "odd_p37_v70_eval"])
def test_agent_act_kernel_on_reference_observations(name):
    """FirstFitAgent.act / BestFitAgent.act on the reference's own observations -> the reference's actions
    (unperturbed fixtures only: the recorded action IS the agent's output)."""
    torch = _torch()
    from vmgym import VecVmEnv
    from vmgym.agents import BestFitAgent, FirstFitAgent
    fx = gu.load(name)
    cfg = fx["cfg"]
    T = fx["action"].shape[0]
    obs = np.concatenate([fx["placement"][:T].astype(np.float64), fx["vm_cpu_code"][:T] / 100.0,
                          fx["vm_mem_code"][:T] / 100.0, fx["cpu"][:T], fx["memory"][:T]], axis=1).astype(np.float32)
    tie = "numpy_introsort" if str(fx["tiebreak"]) == "numpy_introsort" else "stable"
    vec = VecVmEnv(_cfg(**cfg), 1, trace_steps=4, max_admissions=16, tiebreak=tie)
    agent = (FirstFitAgent if str(fx["agent"]) == "firstfit" else BestFitAgent)(vec)
    got = agent.act(torch.from_numpy(obs).to(vec.device)).cpu().numpy().astype(np.int64)
    P, V = cfg["pms"], cfg["vms"]
    otie = vo.TIE_NUMPY_INTROSORT if tie == "numpy_introsort" else vo.TIE_STABLE
    for t in range(T):
        want = vo.firstfit_act(P, V, obs[t]) if str(fx["agent"]) == "firstfit" else vo.bestfit_act(P, V, obs[t], otie)
        assert np.array_equal(got[t], want), t
    if name != "odd_p37_v70_eval":   # that fixture's recorded actions are perturbed
        assert np.array_equal(got, fx["action"].astype(np.int64))
    # numpy in / numpy out form of the reference API
    one = agent.act(obs[T // 2])
    assert one.dtype == np.int64 and np.array_equal(one, got[T // 2])


def test_invalid_action_mask_matches_oracle():
    from vmgym import VecVmEnv
    kw = SHAPES["s100"]
    n_envs, steps = 5, 700
    seeds = 40 + np.arange(n_envs)
    vec = VecVmEnv(_cfg(**kw), n_envs, seeds=seeds, trace_steps=steps + 4, max_admissions=20000)
    oracles = _oracle_batch(kw, seeds, steps + 4, 20000)
    for t in range(0, steps, 100):
        vec.agent_step("firstfit", n_steps=100, want_obs=False)
        for o in oracles:
            o.rollout(vo.AGENT_FIRSTFIT, 100)
        m = vec.get_invalid_action_mask(True).cpu().numpy()
        for i, o in enumerate(oracles):
            assert np.array_equal(m[i], o.get_invalid_action_mask(True)), (i, t)
    assert not vec.get_invalid_action_mask(False).any()


@pytest.mark.parametrize("shape,n_envs,steps,agent", [("s100", 6, 900, "firstfit"), ("s10", 12, 390, "firstfit"),
                                                       ("wide", 3, 300, "firstfit"), ("s100", 4, 4200, "bestfit"),
                                                       ("s10wr", 12, 390, "firstfit"), ("s10wr", 12, 390, "bestfit"),
                                                       # reward wr + Philox + stable ties on u16 placements: the specialised team-mode kernels
                                                       ("big", 3, 300, "bestfit"), ("big", 3, 300, "firstfit"), ("s1000", 2, 1100, "bestfit")])
def test_philox_mode_matches_oracle_on_same_draws(shape, n_envs, steps, agent):
    """rng='philox': the kernel's in-flight Philox/inverse-CDF draws == the host restatement of the same
    counters fed to the oracle env as a pre-sampled trace (arrivals, sizes, service lengths, cursors)."""
    from vmgym import VecVmEnv
    kw = SHAPES[shape]
    seeds = np.array([3, 2**31 + 5, 2**40 + 1, 17, 99, 12345678901, 8, 21, 34, 55, 89, 144][:n_envs], dtype=np.int64)
    vec = VecVmEnv(_cfg(**kw), n_envs, rng="philox", seeds=seeds)
    ka, ta, ks, ts, lo, hi = vec.philox_tables
    oracles = []
    for s in seeds:
        o = vo.OracleVmEnv(vo.OracleConfig(**dict(kw, seed=0)), trace_steps=4, trace_adm=4)
        o.reset(trace=vo.philox_trace(int(s), steps + 4, 40000, ka, ta, ks, ts, lo, hi))
        oracles.append(o)
    P, V = kw["pms"], kw["vms"]
    chunk = 130 if steps < 2000 else 700
    for t0 in range(0, steps, chunk):
        n = min(chunk, steps - t0)
        vec.agent_step(agent, n_steps=n, want_action=False, want_valid=False)
        for o in oracles:
            o.rollout(vo.AGENT_FIRSTFIT if agent == "firstfit" else vo.AGENT_BESTFIT, n)
        _compare_state(vec, oracles, t0 + n)
    obs = vec.obs.cpu().numpy()
    for i, o in enumerate(oracles):
        assert obs[i].tobytes() == o._obs().tobytes()
    del P, V


@pytest.mark.parametrize("shape,N,T", [("s100", 4096, 1500), ("s1000", 1024, 1300)])
def test_full_size_invariants(shape, N, T):
    """BASELINE config 2 size (4096 envs x 100 PMs) and config 5 size (1024 envs x 1000 PMs, team-mode kernel, past the first
    departure wave): size-independent properties after a long fused rollout."""
    from vmgym import VecVmEnv
    kw = dict(SHAPES[shape])
    vec = VecVmEnv(_cfg(**kw), N, rng="philox")
    vec.agent_step("bestfit", n_steps=T, want_obs=True)
    P, V = kw["pms"], kw["vms"]
    place = vec.vm_placement.cpu().numpy().astype(np.int64)
    cc = (vec.vm_cpu_code.cpu().numpy() & 0x7f).astype(np.int64)
    mc = vec.vm_mem_code.cpu().numpy().astype(np.int64)
    cpu, mem = vec.cpu.cpu().numpy(), vec.memory.cpu().numpy()
    cnt = vec.counters()
    # conservation: every request is dropped, served, or still in a slot
    existing = (place <= P).sum(1)
    assert np.array_equal(cnt["total_requests"], cnt["dropped_requests"] + cnt["served_requests"] + existing)
    assert np.all(cnt["timestep"] == T + 1) and np.all(cnt["arrival_pos"] == T)
    assert cnt["served_requests"].min() > 0 and cnt["place_actions"].min() > 0
    assert np.array_equal(cnt["admission_pos"], cnt["served_requests"] + existing)
    # PM accumulators equal the sum of the sizes placed on them (up to fp64 drift), and never exceed capacity
    onehot_cpu = np.zeros((N, P + 2)); onehot_mem = np.zeros((N, P + 2))
    rows = np.repeat(np.arange(N), V)
    np.add.at(onehot_cpu, (rows, place.reshape(-1)), cc.reshape(-1) / 100.0)
    np.add.at(onehot_mem, (rows, place.reshape(-1)), mc.reshape(-1) / 100.0)
    assert np.allclose(cpu, onehot_cpu[:, :P], atol=1e-9) and np.allclose(mem, onehot_mem[:, :P], atol=1e-9)
    assert cpu.max() <= 1.0 and mem.max() <= 1.0 and cpu.min() >= 0.0
    # empty slots carry no size / runtime; running and waiting VMs have a positive remaining runtime
    rem = vec.vm_remaining_runtime.cpu().numpy().astype(np.int64)
    assert np.all(cc[place == P + 1] == 0) and np.all(rem[place == P + 1] == 0)
    assert np.all(rem[place <= P] > 0) and np.all(cc[place <= P] >= (25 if kw.get("sequence") == "highuniform" else 10))
    # the observation is the float32 image of the state
    obs = vec.obs.cpu().numpy()
    want = np.concatenate([place, cc / 100.0, mc / 100.0, cpu, mem], axis=1).astype(np.float32)
    assert obs.tobytes() == want.tobytes()
    # envs are independent: a different batch composition reproduces env 7's trajectory exactly
    vec2 = VecVmEnv(_cfg(**kw), 3, rng="philox", seeds=[kw.get("seed", 0) + 7, 12345, 999])
    vec2.agent_step("bestfit", n_steps=T, want_obs=True)
    assert vec2.state[0].cpu().numpy()[: vec2._layout.off_scalars].tobytes() == \
        vec.state[7].cpu().numpy()[: vec._layout.off_scalars].tobytes()


# ---- the reference's published rows, reproduced by the CUDA path (5 seeds x 100 000 eval steps, fused agents) ----------
GPU_KATS = {
    # name: (config base, arrival_rate, reward, agent, tiebreak, seeds, per-seed returns, csv row) — SURVEY §8c KAT-1..4;
    # arrival_rate = round(pms / 0.55 / service_length, 4) (exp_performance.py:26, exp_performance_small.py:23)
    "KAT1_firstfit_s10": ("10", 0.0182, "ut", "firstfit", "stable", [1, 2, 3, 4, 5],
                          [702661.695, 690616.030, 695187.965, 695977.240, 700256.885],
                          dict(ret=696939.963, drop=0.241, served=1258, cpu=0.697, var=0.051, mem=0.697, wait=0.539, pend=0.183, slow=0.000)),
    "KAT2_firstfit_s100": ("100", 0.1818, "wr", "firstfit", "stable", [0, 1, 2, 3, 4],
                           [-53615.794, -53208.140, -53318.037, -52963.827, -53735.479],
                           dict(ret=-53368.255, drop=0.203, served=13394, cpu=0.737, var=0.052, mem=0.736, wait=0.534, pend=0.070, slow=0.000)),
    "KAT3_bestfit_s10": ("10", 0.0182, "ut", "bestfit", "stable", [1, 2, 3, 4, 5],
                         [699859.415, 694840.720, 698802.250, 697559.640, 703166.045],
                         dict(ret=698845.614, drop=0.242, served=1260, cpu=0.699, var=0.053, mem=0.699, wait=0.537, pend=0.178, slow=0.000)),
    "KAT4_bestfit_s100": ("100", 0.1818, "wr", "bestfit", "numpy_introsort", [0, 1, 2, 3, 4],
                          [-52064.658, -51690.685, -51570.904, -51202.317, -51729.035],
                          dict(ret=-51651.520, drop=0.182, served=13862, cpu=0.763, var=0.057, mem=0.762, wait=0.517, pend=0.068, slow=0.000)),
}


@pytest.mark.parametrize("kat", sorted(GPU_KATS))
def test_published_rows_on_the_cuda_path(kat):
    """data/exp_performance_small/summary.csv:3-4 and data/exp_performance/summary.csv:3-4 of the reference, every printed
    digit, from VecVmEnv.evaluate (numpy-exact traces, fused agent kernel, on-device episode statistics)."""
    from vmgym import VecVmEnv
    base, lam, reward, agent, tie, seeds, per_seed, row = GPU_KATS[kat]
    P, V = (10, 30) if base == "10" else (100, 300)
    cfg = _cfg(pms=P, vms=V, service_length=1000, arrival_rate=lam, training_steps=10000, eval_steps=100000,
               reward_function=reward, cap_target_util=True, sequence="uniform", beta=0.5, allow_null_action=True)
    vec = VecVmEnv(cfg, len(seeds), seeds=seeds, tiebreak=tie).enable_vm_stats()
    s = vec.evaluate(agent, seeds=seeds)
    assert np.all(s["steps"] == 100000)
    for got, want in zip(s["total rewards"], per_seed):
        assert round(float(got), 3) == pytest.approx(want, abs=2e-3)
    assert "%.3f" % s["total rewards"].mean() == "%.3f" % row["ret"]
    assert "%.3f" % s["drop rate"].mean() == "%.3f" % row["drop"]
    assert "%d" % s["total served VMs"].mean() == "%d" % row["served"]
    assert "%.3f" % s["cpu mean"].mean() == "%.3f" % row["cpu"]
    assert "%.3f" % s["cpu var"].mean() == "%.3f" % row["var"]
    assert "%.3f" % s["memory mean"].mean() == "%.3f" % row["mem"]
    assert "%.3f" % s["waiting ratio"].mean() == "%.3f" % row["wait"]
    # Pending Rate / Slowdown Rate columns (exp_performance.py:104-105,139-141): mean over seeds of np.mean(record lists)
    assert "%.3f" % s["average pending"].mean() == "%.3f" % row["pend"]
    assert "%.3f" % s["average slowdown"].mean() == "%.3f" % row["slow"]


@pytest.mark.parametrize("name", ["rec_busy_firstfit", "rec_busy_suspend", "rec_p37_v70", "rec_s10_sparse", "rec_s100_bestfit"])
def test_per_vm_statistics_match_reference_record(name):
    """Record's per-VM lists (record.py:34-96) from the kernel's four clocks per slot + histograms == the lists the
    reference's own Record produced for the same episode (tests/golden/record.npz): same multiset of pending rates,
    slowdown rates and lifetimes, hence the same mean / median / max."""
    import json
    import os
    from vmgym import VecVmEnv
    from vmgym import _native as nv
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "record.npz"))
    cfgd = json.loads(str(z[f"{name}.cfg_json"]))
    actions, pending, slowdown, life = (z[f"{name}.{k}"] for k in ("actions", "pending", "slowdown", "lifetime"))
    summary = json.loads(str(z[f"{name}.summary_json"]))
    T = actions.shape[0]
    vec = VecVmEnv(_cfg(**cfgd), 1, seeds=[cfgd["seed"]], trace_steps=T + 8, max_admissions=4 * T + 512).enable_vm_stats()
    vec.eval(True)
    vec.reset(seed=np.array([cfgd["seed"]]))
    for t in range(T):
        _, _, term, _, _ = vec.step(actions[t:t + 1].astype(np.int64))
    assert bool(term[0])
    s = vec.vm_stats()
    assert int(s["vms"][0]) == pending.size
    assert int(s["placed vms"][0]) in ((slowdown.size,) if slowdown.size != 1 else (0, 1))     # record.py:83-84: [0] when none
    # the histograms are the multisets of the golden lists
    torch = _torch()
    hist = torch.empty_like(vec._vm_hist); totals = torch.empty_like(vec._vm_totals)
    import ctypes as C
    nv.check(nv.lib().vmgym_vmstats_finalize(C.byref(vec._ccfg()), vec.state.data_ptr(), 1, vec._vm_slots.data_ptr(),
                                             vec._vm_hist.data_ptr(), vec._vm_totals.data_ptr(), hist.data_ptr(), totals.data_ptr(),
                                             vec._stream()), "finalize")
    h = hist.cpu().numpy()[0]
    want_p = np.bincount(np.rint(pending * 1000).astype(np.int64), minlength=nv.VMSTAT_BINS)
    assert np.array_equal(h[0], want_p)
    if int(s["placed vms"][0]) > 0:
        want_s = np.bincount(np.rint(slowdown * 1000).astype(np.int64), minlength=nv.VMSTAT_BINS)
        assert np.array_equal(h[1], want_s)
    assert int(totals[0, 2]) == int(life.sum())
    for k in ("average VM life", "average pending", "median pending", "max pending", "average slowdown", "median slowdown",
              "max slowdown"):
        assert float(np.round(s[k][0], 3)) == summary[k], k        # np.round as in record.py:118-125


@pytest.mark.parametrize("P,V,rate,N,T", [(10, 30, 0.45, 6, 700), (260, 600, 9.0, 3, 260)])
def test_per_vm_statistics_batch_vs_oracle_record(P, V, rate, N, T):
    """Philox traces, random suspend / re-place actions mixed into first-fit (no episode end: the statistics include the
    VMs still in their slots) == the oracle's Record restatement on the same traces.  Second case: u16 placements, i.e. the
    team-mode kernel with per-VM clocks."""
    torch = _torch()
    import vmoracle as vo
    from vmgym import VecVmEnv
    from vmgym.agents import FirstFitAgent
    kw = dict(pms=P, vms=V, arrival_rate=rate, service_length=35, training_steps=5000, eval_steps=5000, reward_function="wr",
              allow_null_action=True)
    vec = VecVmEnv(_cfg(**kw), N, rng="philox").enable_vm_stats()
    vec.reset(seed=100 + np.arange(N))
    agent = FirstFitAgent(vec)
    g = torch.Generator(device="cpu").manual_seed(3)
    ka, ta, ks, ts, lo, hi = vec.philox_tables
    oracles = []
    for i in range(N):
        o = vo.OracleVmEnv(vo.OracleConfig(**kw))
        o.enable_record(T, max_arrivals=4 * T + V + int(2 * rate * T))
        o.reset(trace=vo.philox_trace(100 + i, T + 8, 4 * T + 64 + int(2 * rate * T), ka, ta, ks, ts, lo, hi))
        oracles.append(o)
    obs = vec.observe()
    for t in range(T):
        act = agent.act(obs).to(torch.int64)
        u = torch.rand(act.shape, generator=g).to(act.device)
        act = torch.where(u < 0.08, torch.full_like(act, kw["pms"]), act)
        obs, _, _, _, _ = vec.step(act)
        a = act.cpu().numpy()
        for i, o in enumerate(oracles):
            o.step(a[i])
    s = vec.vm_stats()
    for i, o in enumerate(oracles):
        want = o.record_summary()
        pending, slowdown, life = o.record_lists()
        assert int(s["vms"][i]) == len(pending)
        for k, v in want.items():
            assert float(np.round(s[k][i], 3)) == float(v), (i, k)


def test_empty_batch_and_bad_arguments():
    """Zero envs is a no-op, not an error; API misuse returns a status instead of launching (no exception crosses the C ABI)."""
    import ctypes as C
    torch = _torch()
    from vmgym import VecVmEnv
    from vmgym import _native as nv
    vec = VecVmEnv(_cfg(**SHAPES["s10"]), 0, rng="philox")
    assert vec.obs.shape == (0, 110)
    vec.agent_step("firstfit", n_steps=3)
    obs, r, term, trunc, info = vec.step(torch.zeros((0, 30), dtype=torch.int64, device=vec.device))
    assert obs.shape == (0, 110) and r.numel() == 0 and vec.get_invalid_action_mask().shape == (0, 30, 12)
    vec1 = VecVmEnv(_cfg(**SHAPES["s10"]), 2, rng="philox")
    with pytest.raises(ValueError):
        vec1.step(torch.zeros((2, 29), dtype=torch.int64, device=vec1.device))
    lib = nv.lib()
    out = nv.Outputs()
    rc = lib.vmgym_step(C.byref(vec1._ccfg()), vec1.state.data_ptr(), 2, C.byref(vec1._trace), None, nv.I64, C.byref(out), None)
    assert rc == nv.EINVAL and b"action" in lib.vmgym_last_error()
    rc = lib.vmgym_agent_step(C.byref(vec1._ccfg()), vec1.state.data_ptr(), 2, C.byref(vec1._trace), 9, 0, 1, C.byref(out), None)
    assert rc == nv.EUNSUPPORTED
    torch.cuda.synchronize()


@pytest.mark.parametrize("shape", ["s10wr", "s100", "odd", "p254", "wide"])
@pytest.mark.parametrize("agent,tiebreak", [("firstfit", None), ("bestfit", "stable"), ("bestfit", "numpy_introsort")])
def test_step_next_action_equals_agent_act_on_the_new_observation(shape, agent, tiebreak):
    """vmgym_outputs.d_next_action: after a step on external actions the kernel also leaves the heuristic agent's act() on the NEW
    state; it must equal the stand-alone agent.act(obs) on the returned observation (firstfit.py:21-38 / bestfit.py:21-40), also
    when the output tensor is the action tensor itself."""
    import torch
    from vmgym import VecVmEnv
    from vmgym.agents import BestFitAgent, FirstFitAgent
    N = 6 if shape == "wide" else 40
    vec = VecVmEnv(_cfg(**SHAPES[shape]), N, rng="philox", tiebreak=tiebreak or "stable")
    ag = (FirstFitAgent if agent == "firstfit" else BestFitAgent)(vec, tiebreak=tiebreak)
    obs = vec.observe()
    act = ag.act(obs).clone().to(vec.place_dtype)
    nxt = torch.empty_like(act)
    for k in range(60):
        in_place = k % 2 == 1
        out = act if in_place else nxt
        obs, *_ = vec.step(act, want_valid=False, next_action=(agent, out, tiebreak))
        want = ag.act(obs).to(vec.place_dtype)
        torch.cuda.synchronize()
        assert torch.equal(out, want), f"step {k}: next_action differs from agent.act(obs)"
        act = want.clone()
    assert int(vec.counters()["place_actions"].sum()) > 0


def test_obs_mirror_update_stores_exactly_the_changed_entries():
    """vmgym_obs_mirror_update: host copy and device shadow follow the device observations; untouched entries are not written
    (a poisoned host value under an unchanged entry survives), -0.0 vs 0.0 counts as a change (bit compare); reward / done forwarded."""
    import ctypes as C
    import torch
    from vmgym import _native as nv
    N, D = 37, 1100
    lib = nv.lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.rand((N, D), device="cuda", generator=g)
    shadow = obs.clone()
    host = obs.cpu().pin_memory()
    rew_d = torch.rand(N, device="cuda", dtype=torch.float64, generator=g)
    term_d = (torch.rand(N, device="cuda", generator=g) < 0.3).to(torch.uint8)
    rew_h = torch.zeros(N, dtype=torch.float64).pin_memory()
    term_h = torch.zeros(N, dtype=torch.uint8).pin_memory()
    idx = torch.randint(0, N * D, (500,), device="cuda", generator=g).unique()
    flat = obs.view(-1)
    flat[idx] = flat[idx] + 1.0
    flat[7] = -0.0 if float(flat[7]) == 0.0 else flat[7]
    shadow.view(-1)[11] = 0.0
    flat[11] = -0.0                                     # same value, different bits
    host.view(-1)[5] = 123.0 if 5 not in idx.tolist() and 5 not in (7, 11) else host.view(-1)[5]
    poisoned = float(host.view(-1)[5]) == 123.0
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nv.check(lib.vmgym_obs_mirror_update(obs.data_ptr(), shadow.data_ptr(), host.data_ptr(), N * D, rew_d.data_ptr(), rew_h.data_ptr(),
                                         term_d.data_ptr(), term_h.data_ptr(), N, st), "mirror")
    torch.cuda.synchronize()
    assert torch.equal(shadow.view(torch.int32), obs.view(torch.int32))
    want = obs.cpu()
    if poisoned:
        want.view(-1)[5] = 123.0                       # never rewritten: the entry did not change
    assert torch.equal(host.view(torch.int32), want.view(torch.int32))
    assert torch.equal(rew_h, rew_d.cpu()) and torch.equal(term_h, term_d.cpu())
    rc = lib.vmgym_obs_mirror_update(obs.data_ptr(), shadow.data_ptr(), host.data_ptr(), N * D + 2, None, None, None, None, 0, st)
    assert rc == nv.EINVAL


@pytest.mark.parametrize("use_graphs,zero_copy,delta_obs,resident", [(True, True, True, True), (False, True, True, True), (True, True, False, True),
                                                                     (True, False, False, True), (True, True, True, False), (False, False, False, False)])
def test_host_vec_env_pipelined_equals_fused_device_rollout(use_graphs, zero_copy, delta_obs, resident):
    """HostVecEnv (host obs/action buffers, 3 groups on 3 streams, split-phase pipelining) runs the same envs as one
    VecVmEnv stepping the fused best-fit kernel: identical observations, rewards and counters after 60 steps."""
    import torch
    from vmgym import Config, VecVmEnv
    from vmgym.host_vec import HostVecEnv
    cfg = Config(pms=100, vms=300, arrival_rate=1.8182, service_length=1000, training_steps=10000, eval_steps=100000,
                 reward_function="wr", allow_null_action=True)
    N = 50
    hv = HostVecEnv(cfg, N, groups=3, agent="bestfit", use_graphs=use_graphs, zero_copy=zero_copy, delta_obs=delta_obs, resident_obs=resident)
    assert hv.h2d_bytes_per_step == N * 300 + (0 if resident else N * 1100 * 4)
    obs0 = hv.reset().clone()
    ref = VecVmEnv(cfg, N, rng="philox")
    assert torch.equal(obs0, ref.observe().cpu())
    # plain loop for 25 steps, pipelined for 35 (free-running scheduler, then blocking round-robin)
    for _ in range(25):
        hv.act()
        hv.step()
    hv.run_pipelined(20)
    obs, rew, term = hv.run_pipelined(15, poll=False)
    robs, rrew, rterm = ref.agent_step("bestfit", 60, want_obs=True)
    torch.cuda.synchronize()
    assert torch.equal(obs, robs.cpu())
    assert torch.equal(rew, rrew.cpu()) and torch.equal(term.bool(), rterm.cpu())
    c, rc = hv.counters(), ref.counters()
    for k in rc:
        if k != "status":                              # the fused kernel's QUIET flag is not part of the env's state
            assert np.array_equal(np.asarray(c[k]), np.asarray(rc[k])), k
    assert c["place_actions"].sum() > 0
    # act() on a caller's own observation array uploads it and leaves the env's buffers alone
    mine = hv.obs.clone().numpy()
    a_env = hv.act().clone()
    a_mine = hv.act(mine).clone()
    assert torch.equal(a_env, a_mine)
    mine[:, 3 * hv.V:] = 1.0                           # every PM full: best-fit proposes nothing, all actions = current placement
    a_full = hv.act(mine)
    assert torch.equal(a_full.long(), hv.obs[:, :hv.V].long()) and torch.equal(hv.obs, robs.cpu())
    # explicit host actions: an all-WAIT/NULL-preserving no-op action leaves placements unchanged
    act = hv.obs[:, :hv.V].to(hv.place_dtype).numpy()
    o2, _, _ = hv.step(act)
    ref.step(ref.vm_placement.clone())
    assert torch.equal(o2, ref.observe().cpu())
    # fast_forward with one step count per group (the benchmark's phase staggering) == the same counts on env slices
    hv.fast_forward([3, 0, 5])
    for gi, n in enumerate([3, 0, 5]):
        g = hv.groups[gi]
        if n:
            ref.agent_step("bestfit", n, want_obs=True, envs=(g.lo, g.hi))
    torch.cuda.synchronize()
    assert torch.equal(hv.obs, ref.observe().cpu())
    with pytest.raises(ValueError):
        hv.fast_forward([1, 2])


# ---- the reference's sweep drivers as one batch per agent (vmgym.sweep) --------------------------------------------
S100_BASE = dict(pms=100, vms=300, service_length=1000, arrival_rate=1.8182, training_steps=10000, eval_steps=100000, seed=0,
                 reward_function="kl", cap_target_util=True, sequence="uniform", beta=0.5, allow_null_action=True)   # config/100.yml
SUSPENSION_ROWS = [          # data/exp_suspension/data.csv of the reference, first-fit / best-fit rows
    "bestfit,0.5,1000,9123,0,9223,993,0.002,0.000,0.000", "bestfit,0.6,1000,10879,0,10977,994,0.003,0.000,0.000",
    "bestfit,0.7,1000,12592,0,12718,994,0.009,0.000,0.000", "bestfit,0.8,1000,13808,0,13943,983,0.073,0.000,0.000",
    "bestfit,0.9,1000,13803,0,13937,983,0.069,0.000,0.000", "bestfit,1.0,100,137105,0,137243,98,0.062,0.000,0.000",
    "bestfit,1.0,1000,13802,0,13935,983,0.072,0.000,0.000", "bestfit,1.0,1100,12523,0,12664,1080,0.072,0.000,0.000",
    "bestfit,1.0,2100,6588,0,6731,2030,0.087,0.000,0.000", "bestfit,1.0,3100,4466,0,4603,2948,0.094,0.000,0.000",
    "bestfit,1.0,4100,3376,0,3512,3840,0.107,0.000,0.000", "bestfit,1.1,1000,13853,0,13991,982,0.073,0.000,0.000",
    "firstfit,0.5,1000,9123,0,9223,993,0.002,0.000,0.000", "firstfit,0.6,1000,10878,0,10976,994,0.004,0.000,0.000",
    "firstfit,0.7,1000,12592,0,12716,994,0.025,0.000,0.000", "firstfit,0.8,1000,13362,0,13499,982,0.071,0.000,0.000",
    "firstfit,0.9,1000,13368,0,13505,983,0.069,0.000,0.000", "firstfit,1.0,100,130390,0,130515,98,0.064,0.000,0.000",
    "firstfit,1.0,1000,13367,0,13509,982,0.076,0.000,0.000", "firstfit,1.0,1100,12180,0,12313,1079,0.072,0.000,0.000",
    "firstfit,1.0,2100,6404,0,6533,2028,0.095,0.000,0.000", "firstfit,1.0,3100,4332,0,4464,2944,0.101,0.000,0.000",
    "firstfit,1.0,4100,3296,0,3422,3834,0.111,0.000,0.000", "firstfit,1.1,1000,13388,0,13527,982,0.069,0.000,0.000",
]


@pytest.mark.parametrize("agent", ["firstfit", "bestfit"])
def test_suspension_sweep_reproduces_published_table(agent):
    """exp_suspension.py's (load, service length) grid as ONE 12-env batch per agent: every printed digit of the 24
    first-fit / best-fit rows of data/exp_suspension/data.csv (served, valid actions, mean life, pending, slowdown)."""
    from vmgym.sweep import suspension_points, suspension_rows
    pts = suspension_points(100, service_lengths=np.arange(100, 4200, 1000), loads=np.arange(0.5, 1.15, 0.1))
    rows = suspension_rows(S100_BASE, agent, points=pts)
    want = sorted(r for r in SUSPENSION_ROWS if r.startswith(agent))
    assert sorted(rows) == want


VM_SIZE_ROWS = {             # data/exp_vm_size/summary.csv of the reference (reward kl, seeds 0..4): lowuniform, highuniform
    "firstfit": ["firstfit,22539.4184,0.1232,22602,0,0.8504,0.0179,0.8497,0.0180,0.2226",
                 "firstfit,-73517.7000,0.2224,11454,0,0.7175,0.0435,0.7174,0.0437,0.5992"],
    "bestfit": ["bestfit,51673.7852,0.1093,23057,0,0.8677,0.0189,0.8672,0.0190,0.2056",
                "bestfit,-56063.0182,0.2159,11600,0,0.7267,0.0467,0.7266,0.0467,0.5939"],
}


@pytest.mark.parametrize("agent", ["firstfit", "bestfit"])
def test_vm_size_sweep_reproduces_published_table(agent):
    """exp_vm_size.py (lowuniform / highuniform size mixes, reward kl, 5 seeds averaged) through vmgym.sweep: all columns
    as printed; the kl return (a sum of 500 000 log-containing rewards) to 2e-3 absolute."""
    from vmgym.sweep import vm_size_rows
    rows = vm_size_rows(S100_BASE, agent)
    for got, want in zip(rows, VM_SIZE_ROWS[agent]):
        g, w = got.split(","), want.split(",")
        assert g[0] == w[0] and g[2:] == w[2:], (got, want)
        assert abs(float(g[1]) - float(w[1])) <= 2e-3, (got, want)


@pytest.mark.parametrize("case,agent,reward", [("main_s10_firstfit_wr", "firstfit", "wr"), ("main_s10_bestfit_ut", "bestfit", "ut")])
def test_main_entry_reproduces_reference_summary(case, agent, reward, tmp_path):
    """`python -m vmgym.main -a <agent> -r <reward> -e -c configs/10.yml` (the reference's main.py:31-87 flow: seed, env, agent,
    Base.test, Record) prints the summary the reference's own Record produced for the same command (tests/golden/record.npz,
    100 000 steps, real SVD rank): all 22 keys of Record.get_summary, as rounded by the reference."""
    import json
    import os
    import yaml
    from vmgym.main import Args, run
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "record.npz"))
    want = json.loads(str(z[f"{case}.summary_json"]))
    cfg = yaml.safe_load(open(os.path.join(root, "configs", "10.yml")))
    out = tmp_path / "record.json"
    rec = run(Args(agent=agent, reward=reward, config=cfg, eval=True, silent=True, output=str(out)))
    got = rec.get_summary()
    assert list(got) == list(want)                               # same keys, same order
    for k, v in want.items():
        assert got[k] == pytest.approx(v, abs=1e-9 if isinstance(got[k], int) else 5e-4), (k, got[k], v)
    exact = sum(1 for k, v in want.items() if got[k] == v)
    assert exact >= len(want) - 1, {k: (got[k], v) for k, v in want.items() if got[k] != v}      # at most one last-digit flip
    assert json.load(open(out))["summary"] == got


def test_bench_gpu_arm_prints_the_contract_line():
    """`python bench.py --steps 6 --warmup 3 --no-cpu --no-extras`: one JSON line with every key of the bench contract."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "6", "--warmup", "3", "--no-cpu", "--no-extras",
                          "--batches", "4", "--envs", "512", "--replays", "5"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "gpu_launches", "e2e", "roofline", "clocks"):
        assert k in d, k
    # gpu_launches: the timed region is 5 replays of ONE persistent launch of 6 batch steps (vmgym_agent_step_rotation)
    assert d["steps"] == 6 and d["warmup"] == 3 and d["gpu_launches"] == 5 and d["n_gpus"] == 1 and d["value"] > 0
    ts = d["timing_stats"]
    assert ts["replays"] == 5 and ts["steps_per_replay"] == 6 and ts["p10_ms"] <= ts["median_ms"] <= ts["p90_ms"]
    assert abs(d["ms_per_step"] - ts["median_ms"] / 6) < 1e-12
    r = d["roofline"]
    assert r["B_fused"] == 2 * (16 * 100 + 5 * 300 + 48) + 16 and 0.0 <= r["obs_rows_stored_frac"] <= 1.0
    assert abs(r["bytes_per_env_step"] - (r["B_fused"] + r["obs_bytes_stored_per_env_step"])) < 1e-6
    assert d["unit"] == "env-steps/s" and d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9 and "workload" in d["config"]


PERFORMANCE_ROWS = [     # data/exp_performance/summary.csv (S100, reward ut at load 0.60 / wr at 1.00) and data/exp_performance_small/summary.csv
    ("100", "firstfit", 0.60, "ut", range(5), "firstfit,0.60,5979819.544,0.000,10825,0,0.598,0.105,0.598,0.066,0.003,0.002,0.000"),
    ("100", "bestfit", 0.60, "ut", range(5), "bestfit,0.60,5979835.178,0.000,10825,0,0.598,0.131,0.598,0.102,0.003,0.002,0.000"),
    ("100", "bestfit", 1.00, "wr", range(5), "bestfit,1.00,-51651.520,0.182,13862,0,0.763,0.057,0.762,0.047,0.068,0.517,0.000"),
    ("10", "firstfit", 1.00, "ut", range(1, 6), "firstfit,1.00,696939.963,0.241,1258,0,0.697,0.051,0.697,0.046,0.183,0.539,0.000"),
]


@pytest.mark.parametrize("shape,agent,load,reward,seeds,want", PERFORMANCE_ROWS)
def test_performance_rows_through_sweep(shape, agent, load, reward, seeds, want):
    """exp_performance.py's table rows through vmgym.sweep.performance_row (5 seeds = one 5-env batch): every column as
    printed except Memory Variance, which the reference computes ACROSS THE RUNS (np.var(memory, axis=0), :117)."""
    from vmgym.sweep import performance_row
    base = dict(S100_BASE) if shape == "100" else dict(S100_BASE, pms=10, vms=30, seed=1)
    got = performance_row(base, agent, load, reward, seeds=seeds).split(",")
    w = want.split(",")
    assert got[:9] == w[:9] and got[10:] == w[10:], (",".join(got), want)


@pytest.mark.parametrize("shape,E,NB,K,n_steps,agent,bulk", [("s100", 300, 3, 8, 1, "bestfit", 7), ("s100", 4200, 2, 5, 1, "bestfit", 7),
                                                            ("s10", 700, 4, 9, 2, "firstfit", 7 | 32), ("wide", 5, 3, 7, 1, "firstfit", 7),
                                                            ("s100", 64, 5, 12, 3, "firstfit", 0),
                                                            # team mode with more envs per batch than resident CTAs: records are shared out
                                                            # as (batch * E + index) mod grid (balanced team rotation)
                                                            ("p254", 5000, 2, 5, 1, "firstfit", 7), ("big", 1300, 3, 7, 2, "bestfit", 7),
                                                            ("s1000", 700, 2, 4, 1, "bestfit", 7)])
def test_rotation_launch_equals_per_batch_launches(shape, E, NB, K, n_steps, agent, bulk):
    """vmgym_agent_step_rotation: ONE persistent launch over NB sub-batches x K batch steps == K calls of the fused agent step on
    the sub-batches in rotation (byte-identical records, observations, rewards, done flags), for record staging by bulk copies,
    double-buffered records, plain loads, the team-mode kernel (wide; p254 / big / s1000 with more envs per batch than resident
    CTAs) and more env indexes than resident warps (4200)."""
    torch = _torch()
    from vmgym import VecVmEnv
    from vmgym import _native as nv
    kw = SHAPES[shape]
    N = E * NB
    seeds = 900 + np.arange(N)
    first = 1 % NB
    nv.lib().vmgym_set_tuning(0, bulk)
    try:
        a = VecVmEnv(_cfg(**kw), N, rng="philox", seeds=seeds)
        b = VecVmEnv(_cfg(**kw), N, rng="philox", seeds=seeds)
        # different phases per sub-batch before the comparison (as the benchmark staggers its batches)
        for vec in (a, b):
            for s in range(NB):
                vec.agent_step(agent, n_steps=30 + 7 * s, want_action=False, want_valid=False, envs=(s * E, (s + 1) * E))
        nxt = a.agent_step_rotation(agent, E, K, first_batch=first, n_steps=n_steps)
        assert nxt == (first + K) % NB
        # reference: the same batch steps as separate launches on b's sub-batches
        for k in range(K):
            s = (first + k) % NB
            b.agent_step(agent, n_steps=n_steps, want_action=False, want_valid=False, envs=(s * E, (s + 1) * E))
        torch.cuda.synchronize()
        L = a._layout
        ra, rb = a.state.cpu().numpy(), b.state.cpu().numpy()
        # everything in the record except the parked Philox words and derived caches must match; compare field by field
        for name, lo, hi in (("cpu", L.off_cpu, L.off_cpu + 8 * a.P), ("memory", L.off_memory, L.off_memory + 8 * a.P),
                             ("remaining", L.off_remaining, L.off_remaining + 2 * a.V),
                             ("placement", L.off_placement, L.off_placement + L.place_bytes * a.V),
                             ("codes", L.off_cpu_code, L.off_cpu_code + a.V), ("mcodes", L.off_mem_code, L.off_mem_code + a.V),
                             ("scalars", L.off_scalars, L.off_scalars + nv.SCALARS_BYTES)):
            assert np.array_equal(ra[:, lo:hi], rb[:, lo:hi]), name
        assert torch.equal(a.obs, b.obs) and torch.equal(a.reward, b.reward) and torch.equal(a.terminated_u8, b.terminated_u8)
        assert int(a.counters()["timestep"].max()) > 31
    finally:
        nv.lib().vmgym_set_tuning(0, 7)
