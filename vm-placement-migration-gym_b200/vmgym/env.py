"""VmEnv — the reference's single-env gym surface (vmenv/envs/env.py:19-325) as a numpy-in / numpy-out facade
over a 1-env VecVmEnv.  Same constructor, attributes and return types, so `Base.test`-style loops
(src/agents/base.py:63-86) and `main.run` (main.py:47) work unchanged."""
from __future__ import annotations

import numpy as np
import torch

from .config import Config
from .vec_env import VecVmEnv


class _Box:
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.float32


class _MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape


class VmEnv:
    metadata = {"render_modes": ["ansi"]}

    def __init__(self, config: Config, device="cuda", rng: str = "numpy", trace_steps=None, max_admissions=None,
                 tiebreak: str = "stable"):
        self.config = config
        self.action_dim = config.action_dim                                   # env.py:26
        self.observation_space = _Box(0, config.pms + 2, (config.obs_dim,))   # env.py:27
        self.action_space = _MultiDiscrete(np.full(config.vms, self.action_dim))   # env.py:28
        self.WAIT_STATUS, self.NULL_STATUS = config.pms, config.pms + 1
        self.vec = VecVmEnv(config, 1, device=device, rng=rng, seeds=[config.seed], trace_steps=trace_steps,
                            max_admissions=max_admissions, tiebreak=tiebreak)
        # the per-slot arrival clocks of the device env feed vm_arrival_steps / vm_planned_runtime (env.py:203,205,289,293);
        # nothing has stepped yet, so the freshly zeroed buffers are consistent with the reset state
        self.vec.enable_vm_stats()
        self._reset_host_logs()
        self.last_validity = self.last_action = self.last_reward = None

    def _reset_host_logs(self):
        V = self.config.vms
        self.vm_planned_runtime = np.zeros(V, dtype=int)                     # env.py:203
        self.vm_arrival_steps = [[] for _ in range(V)]                       # env.py:205

    @property
    def eval_mode(self):
        return self.vec.eval_mode

    def eval(self, eval_mode=True):
        self.vec.eval(eval_mode)

    def seed(self, seed=None):
        self.vec.seed(None if seed is None else [int(seed)])

    def reset(self, seed=None, options=None):
        obs, _ = self.vec.reset(None if seed is None else [int(seed)])
        self._reset_host_logs()
        return obs[0].cpu().numpy(), self._get_info()

    def step(self, action):
        action = np.ascontiguousarray(action, dtype=np.int64).reshape(1, -1)
        vec = self.vec
        obs, reward, term, trunc, info = vec.step(torch.from_numpy(action).to(vec.device))
        D, V, P = vec.obs_dim, vec.V, vec.P
        # one D2H (every value is exactly representable in float64): observation, reward, done, valid flags, the slots'
        # arrival clocks, the step counter and the remaining runtimes (u16 stored in an int16 view)
        out = torch.cat([obs[0].double(), reward[:1], term[:1].double(), vec.valid[0].double(), vec._vm_slots[0, :, 0].double(),
                         vec._scalars_i32[0, :1].double(), vec.vm_remaining_runtime[0].double()]).cpu().numpy()
        obs_np, rew, done = out[:D].astype(np.float32), float(out[D]), bool(out[D + 1])
        valid = out[D + 2:D + 2 + V].astype(np.int64)
        arrived_at = out[D + 2 + V:D + 2 + 2 * V].astype(np.int64)
        t_exec = int(out[D + 2 + 2 * V]) - 1                       # the step just executed (the kernel has advanced the clock)
        rem = out[D + 3 + 2 * V:D + 3 + 3 * V].astype(np.int64)
        place = obs_np[:V].astype(np.int64)
        # env.py:288-293: slots admitted in this step start with remaining == planned and log their arrival at step + 1;
        # env.py:262: a departed slot's planned runtime is zeroed
        for v in np.nonzero(arrived_at == t_exec)[0]:
            self.vm_planned_runtime[v] = rem[v]
            self.vm_arrival_steps[v].append(t_exec + 1)
        self.vm_planned_runtime[place == P + 1] = 0
        info = {"action": action[0].copy(), "valid": valid}
        if self.eval_mode:
            full = self._get_info()
            full["timestep"] -= 1                                   # env.py:165 builds info before `timestep += 1` (:101)
            info = full | info
            self.last_validity, self.last_action = valid, action[0]
            self.last_reward = np.round(rew, 3)
        return obs_np, rew, done, False, info

    def get_invalid_action_mask(self, masked: bool = True):
        return self.vec.get_invalid_action_mask(masked)[0].cpu().numpy()

    def state(self):
        return self.vec.state_dict_host(0)

    def _get_info(self):
        if not self.eval_mode:
            return {}
        s = self.state()
        return {"waiting_ratio": s["waiting_ratio"], "served_requests": s["served_requests"],
                "suspend_actions": s["suspend_actions"], "place_actions": s["place_actions"],
                "dropped_requests": s["dropped_requests"], "total_requests": s["total_requests"],
                "timestep": s["timestep"], "vm_arrival_steps": self.vm_arrival_steps,            # live list, as env.py:307
                "vm_placement": s["vm_placement"], "cpu": s["cpu"], "memory": s["memory"],
                "vm_cpu": s["vm_cpu"], "vm_memory": s["vm_memory"], "target_cpu_mean": s["target_cpu_mean"],
                "target_memory_mean": s["target_memory_mean"], "total_cpu_requested": s["total_cpu_requested"],
                "total_memory_requested": s["total_memory_requested"],
                "rank": int(np.unique(s["vm_placement"][s["vm_placement"] < self.config.pms]).size)}   # env.py:319-325

    # reference attribute names
    def __getattr__(self, name):
        if name in ("vm_placement", "vm_cpu", "vm_memory", "cpu", "memory", "vm_remaining_runtime", "vm_suspended",
                    "timestep", "total_requests", "served_requests", "suspend_action", "place_action",
                    "dropped_requests", "waiting_ratio", "target_cpu_mean", "target_memory_mean",
                    "total_cpu_requested", "total_memory_requested"):
            return self.state()[name]
        raise AttributeError(name)

    def render(self, mode="ansi", close=False):
        s = self.state()
        np.set_printoptions(linewidth=np.inf)
        print(f"Timestep: \t\t{s['timestep']}")
        print(f"VM request: \t\t{int((s['vm_placement'] == self.WAIT_STATUS).sum())}, dropped: {s['dropped_requests']}")
        for label, key in (("VM placement", "vm_placement"), ("VM suspended", "vm_suspended")):
            print(f"{label}: \t\t{s[key]}")
        for label, key in (("CPU (%)", "cpu"), ("Memory (%)", "memory"), ("VM CPU (%)", "vm_cpu"), ("VM Memory (%)", "vm_memory")):
            print(f"{label}: \t\t{np.array(s[key] * 100, dtype=int)} {np.round(np.sum(s[key]), 3)}")
        print(f"VM remaining runtime: \t{s['vm_remaining_runtime']}")

    def close(self):
        pass
