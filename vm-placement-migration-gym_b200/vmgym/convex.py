"""Convex agent bridge (reference src/agents/convex.py:15-187).

The reference's ConvexAgent re-places VMs every `frequency` steps by solving a mixed-integer program with cvxpy + SCIP
(convex.py:80-187) — third-party arithmetic that is neither installed in this image nor part of the accelerated path.  What IS
kept is everything a caller sees:

  * `ConvexAgent(env, config)` with the reference's interface (`act(observation) -> action`, no-op `learn / eval / load_model /
    save_model`) and its action protocol (convex.py:33-77): a solver result is turned into *suspend now, place next step*
    pairs through a queue, because the env allows no direct PM -> PM move (env.py:35-42); the solver is consulted only every
    `frequency` steps (and at the last evaluation step); queued placements are flushed before anything else.
  * the MIP itself behind `solve(P, V, vm_cpu, vm_memory, vm_placement) -> new_placement`: `CvxpyScipSolver` binds the
    reference's own method when the reference and its solver stack are importable; any other callable plugs in.
  * `HostAgentBridge`: host-side agents (anything with `act(obs) -> action`) drive SELECTED envs of a device batch while a fused
    device agent drives the rest — per step only the selected envs' observation rows travel to the host and only their action
    rows travel back.
Parity of the MIP's results is unpinned (solver not available here; the reference publishes a 3-row table only).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _native as nv
from .agents import AgentBase


@dataclass
class ConvexConfig:                       # convex.py:9-13
    W: int = 30
    frequency: int = 3
    timeout: int = 3


class CvxpyScipSolver:
    """`maximize_nuclear_norm` of the reference (convex.py:80-187), bound at call time: needs cvxpy + SCIP and the reference
    package on the path.  Raises VmgymError with the reason otherwise — there is no substitute solver."""

    def __init__(self, config: ConvexConfig):
        self.config = config
        self._impl = None

    def __call__(self, P, V, vm_cpu, vm_memory, vm_placement):
        if self._impl is None:
            try:
                import cvxpy  # noqa: F401
                from src.agents import convex as ref_convex           # the reference module (its own cvxpy program)
                if not hasattr(ref_convex.ConvexAgent, "maximize_nuclear_norm"):
                    raise ImportError("src.agents.convex is not the reference module")
                holder = type("_Holder", (), {"config": self.config})()
                self._impl = lambda *a: ref_convex.ConvexAgent.maximize_nuclear_norm(holder, *a)
            except Exception as e:        # noqa: BLE001
                raise nv.VmgymError(f"ConvexAgent needs the reference's MIP stack (cvxpy + SCIP + src.agents.convex): {e}") from e
        return self._impl(P, V, vm_cpu, vm_memory, vm_placement)


class ConvexAgent(AgentBase):
    """convex.py:15-77 with a pluggable solver.  `env` needs `.config` (pms, vms, eval_steps) and `.timestep`."""

    name = "ConvexAgent"

    def __init__(self, env, config: ConvexConfig | None = None, solve=None):
        self.env = env
        self.config = config or ConvexConfig()
        self.solve = solve or CvxpyScipSolver(self.config)
        self.queue = []                   # (slot, pm) placements that follow a suspension (convex.py:18)
        self.failures = 0

    def eval(self, mode=True):
        pass

    def learn(self):
        pass

    def load_model(self, modelpath):
        pass

    def save_model(self, modelpath):
        pass

    def act(self, observation):
        cfg = self.env.config
        P, V = int(cfg.pms), int(cfg.vms)
        obs = observation.detach().cpu().numpy() if isinstance(observation, torch.Tensor) else np.asarray(observation)
        vm_placement = obs[:V].astype(int)                                 # utils.py:37-48
        vm_cpu, vm_memory = np.array(obs[V:2 * V]), np.array(obs[2 * V:3 * V])
        # 1. pending second halves of migrations first (convex.py:42-45)
        flushed = bool(self.queue)
        for slot, pm in self.queue:
            vm_placement[slot] = pm
        self.queue = []
        # 2. the solver runs every `frequency` steps and at the very last evaluation step (convex.py:47-49)
        t = int(self.env.timestep)
        skip = t % int(self.config.frequency) > 0 and int(cfg.eval_steps) != t
        if flushed or skip:
            return vm_placement
        new_placement = np.asarray(self.solve(P, V, vm_cpu, vm_memory, vm_placement.copy())).astype(int)
        # 3. a VM that moves between two PMs is suspended now and placed in the next step (convex.py:67-75)
        moved = (vm_placement < P) & (new_placement < P) & (vm_placement != new_placement)
        for slot in np.nonzero(moved)[0]:
            self.queue.append((int(slot), int(new_placement[slot])))
            new_placement[slot] = P
        return new_placement


class _EnvView:
    """What a host agent reads of ONE env of a batch: `.config` and `.timestep` (the reference's agents use both)."""

    def __init__(self, vec, index: int):
        self.vec, self.index, self.config = vec, int(index), vec.config
        self.WAIT_STATUS, self.NULL_STATUS, self.action_dim = vec.WAIT_STATUS, vec.NULL_STATUS, vec.action_dim

    @property
    def timestep(self):
        return int(self.vec._scalars_i32[self.index, 0].item())


class HostAgentBridge:
    """Host agents on selected envs of a VecVmEnv, a fused device agent ("firstfit" / "bestfit") on all the others.

        bridge = HostAgentBridge(vec, {3: lambda env: ConvexAgent(env, ConvexConfig()), 17: my_agent_factory}, default="bestfit")
        obs, reward, terminated = bridge.step()

    Per step: the device agent proposes actions for every env (vmgym_agent_act), the observation rows of the selected envs
    are read back (one gather + one D2H), their host agents act, their action rows are scattered over the device actions, and
    ONE env.step advances the whole batch."""

    def __init__(self, vec, agents: dict, default: str = "bestfit"):
        from .agents import BestFitAgent, FirstFitAgent
        self.vec = vec
        self.index = torch.as_tensor(sorted(int(i) for i in agents), dtype=torch.int64, device=vec.device)
        if self.index.numel() and (int(self.index.min()) < 0 or int(self.index.max()) >= vec.num_envs):
            raise ValueError("env index out of range")
        self.agents = [agents[int(i)](_EnvView(vec, int(i))) for i in self.index.tolist()]
        self.default = {"bestfit": BestFitAgent, "firstfit": FirstFitAgent}[default](vec)

    def act(self, obs=None):
        vec = self.vec
        obs = vec.obs if obs is None else obs
        action = self.default.act(obs).clone()
        if self.agents:
            rows = obs.index_select(0, self.index).cpu().numpy()
            host = np.stack([np.asarray(a.act(r)) for a, r in zip(self.agents, rows)]).astype(np.int64)
            action.index_copy_(0, self.index, torch.from_numpy(host).to(vec.device).to(action.dtype))
        return action

    def step(self):
        obs, reward, terminated, _, _ = self.vec.step(self.act())
        return obs, reward, terminated
