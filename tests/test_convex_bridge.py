"""The Convex agent bridge (vmgym/convex.py; reference src/agents/convex.py:15-77): the action protocol on the host (CPU test, a
recorded solver stands in for the MIP) and, -m gpu, host agents driving selected envs of a device batch."""
import types

import numpy as np
import pytest


def _env(P=4, V=6, eval_steps=100, timestep=3):
    return types.SimpleNamespace(config=types.SimpleNamespace(pms=P, vms=V, eval_steps=eval_steps), timestep=timestep)


def _obs(placement, P=4):
    V = len(placement)
    return np.concatenate([np.asarray(placement, np.float32), np.full(V, 0.2, np.float32), np.full(V, 0.3, np.float32),
                           np.zeros(2 * P, np.float32)])


def test_convex_agent_action_protocol():
    from vmgym.convex import ConvexAgent, ConvexConfig
    P = 4
    calls = []

    def solve(P_, V_, cpu, mem, placement):
        calls.append(placement.copy())
        out = placement.copy()
        out[0] = 2            # running on PM 0 -> PM 2: a migration (suspend now, place next step)
        out[1] = 1            # waiting -> PM 1: a plain placement
        return out

    env = _env(timestep=3)
    agent = ConvexAgent(env, ConvexConfig(frequency=3), solve=solve)
    a = agent.act(_obs([0, P, 3, P + 1, P + 1, P + 1]))
    assert a.tolist() == [P, 1, 3, P + 1, P + 1, P + 1] and agent.queue == [(0, 2)] and len(calls) == 1
    assert np.allclose(calls[0], [0, P, 3, P + 1, P + 1, P + 1])
    env.timestep = 4          # the queued half of the migration is flushed first; the solver is not consulted
    a = agent.act(_obs([P, 1, 3, P + 1, P + 1, P + 1]))
    assert a.tolist() == [2, 1, 3, P + 1, P + 1, P + 1] and agent.queue == [] and len(calls) == 1
    env.timestep = 5          # 5 % 3 > 0: skipped
    assert agent.act(_obs([2, 1, 3, P + 1, P + 1, P + 1])).tolist() == [2, 1, 3, P + 1, P + 1, P + 1] and len(calls) == 1
    env.timestep = 100        # == eval_steps: solved although 100 % 3 > 0 (convex.py:47)
    agent.act(_obs([2, 1, 3, P + 1, P + 1, P + 1]))
    assert len(calls) == 2
    for m in ("learn", "eval"):
        getattr(agent, m)()
    agent.load_model("x"); agent.save_model("x")


def test_default_solver_reports_the_missing_mip_stack():
    from vmgym import _native as nv
    from vmgym.convex import ConvexAgent
    agent = ConvexAgent(_env(timestep=3))
    try:
        import cvxpy  # noqa: F401
        pytest.skip("cvxpy is installed here")
    except ImportError:
        pass
    with pytest.raises(nv.VmgymError, match="cvxpy"):
        agent.act(_obs([0, 4, 3, 5, 5, 5]))


def test_reference_import_path():
    from src.agents.convex import ConvexAgent, ConvexConfig
    assert ConvexConfig().W == 30 and ConvexConfig().frequency == 3 and ConvexAgent.name == "ConvexAgent"


@pytest.mark.gpu
def test_host_agents_drive_selected_envs_of_a_batch():
    import torch
    import vmoracle as vo
    from vmgym import Config, VecVmEnv
    from vmgym.convex import ConvexAgent, ConvexConfig, HostAgentBridge
    kw = dict(pms=10, vms=30, arrival_rate=0.4, service_length=25, training_steps=400, eval_steps=1000, reward_function="wr",
              allow_null_action=True)
    N = 12
    a = VecVmEnv(Config(**kw), N, rng="philox")
    b = VecVmEnv(Config(**kw), N, rng="philox")

    class HostFirstFit:               # a host agent: the oracle's restatement of firstfit.py:21-38 on one observation row
        def __init__(self, env):
            self.env = env

        def act(self, obs):
            assert self.env.timestep >= 1
            return vo.firstfit_act(kw["pms"], kw["vms"], np.asarray(obs, np.float32))

    bridge = HostAgentBridge(a, {2: HostFirstFit, 7: HostFirstFit}, default="firstfit")
    for _ in range(120):
        bridge.step()
    b.agent_step("firstfit", n_steps=120, want_action=False, want_valid=False)
    torch.cuda.synchronize()
    assert torch.equal(a.obs, b.obs) and torch.equal(a.vm_placement, b.vm_placement)
    # a ConvexAgent with a recorded solver on env 5: its suspend-then-place protocol goes through env.step of the batch
    def solve(P, V, cpu, mem, placement):
        out = placement.copy()
        running = np.nonzero(placement < P)[0]
        if running.size:
            out[running[0]] = (placement[running[0]] + 1) % P      # ask to move the first running VM to the next PM
        return out
    c = VecVmEnv(Config(**kw), N, rng="philox")
    c.agent_step("firstfit", n_steps=60, want_action=False, want_valid=False)
    br = HostAgentBridge(c, {5: lambda env: ConvexAgent(env, ConvexConfig(frequency=1), solve=solve)}, default="firstfit")
    before = c.counters()["suspend_actions"].copy()
    for _ in range(6):
        br.step()
    after = c.counters()["suspend_actions"]
    assert after[5] > before[5] and (np.delete(after, 5) == np.delete(before, 5)).all()      # only env 5 migrates
