"""Heuristic agents with the reference's interface (src/agents/base.py:15-149, firstfit.py, bestfit.py):
`act(observation) -> action`, plus no-op learn/load_model/save_model/eval.  `act` accepts what the reference
accepts (one float32 observation, numpy) and also a device batch [N, 3V+2P]; the scan itself is the CUDA kernel
`vmgym_agent_act` — agents decide on the float32 view and accumulate locally in float32 exactly like the
reference (SURVEY App. B-2,3)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nv


class AgentBase:
    """The parts of the reference's Base (src/agents/base.py:15-149) every agent shares: optional TensorBoard log,
    and `test` — one evaluation episode (eval mode, reset(seed), act/step until done) returning a Record."""
    name = "Agent"
    writer = None

    def set_log(self, jobname, logdir):                      # base.py:27-41
        if logdir:
            try:
                from time import gmtime, strftime
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(f"{logdir}/{strftime('%Y%m%d', gmtime())}-{self.name}-{jobname}")
            except Exception:
                self.writer = None

    def end_log(self):                                       # base.py:127-129
        if self.writer:
            self.writer.close()

    def _fused_kind(self):
        return None                                          # heuristic agents: the name of the fused device agent

    def test(self, show: bool = False, output: str | None = None, debug: bool = False):
        """base.py:63-124 on every env of the batch.  Heuristic agents run the whole episode in the fused kernel; learned
        agents loop act/step.  The episode statistics (Record.get_summary) are accumulated on the device."""
        from .record import Record
        vec = getattr(self.env, "vec", self.env)
        vec.enable_vm_stats()
        self.eval()
        kind = self._fused_kind()
        if kind is not None and not debug:
            summary = vec.evaluate(kind, seeds=self._test_seeds(vec), tiebreak=getattr(self, "tiebreak", None))
        else:
            vec.eval(True)
            obs, _ = vec.reset(seed=self._test_seeds(vec))
            done = False
            while not done:
                if debug and hasattr(self.env, "render"):
                    self.env.render()
                action = self.act(obs)
                obs, reward, term, _, info = vec.step(action, want_stats=True)
                done = bool(term.all().item())
            summary = vec.summary()
        record = Record(self.name, vec.config, getattr(self, "config", None), summary)
        if show:
            print(vec.config)
            for k, v in record.get_summary().items():
                print("%s: %.2f" % (k, v))
        if output:
            record.save(output)
        self.record = record
        return record

    @staticmethod
    def _test_seeds(vec):
        return np.asarray([int(c.seed) for c in vec.env_configs] if getattr(vec, "env_configs", None)
                          else int(vec.config.seed) + np.arange(vec.num_envs), np.int64)


class HeuristicAgent(AgentBase):
    kind = nv.AGENT_NONE
    name = "HeuristicAgent"

    def __init__(self, env, tiebreak: str | None = None):
        self.env = env
        self.vec = getattr(env, "vec", env)            # VmEnv facade or VecVmEnv
        self.config = None
        self.tiebreak = tiebreak or getattr(self.vec, "tiebreak", "stable")
        self.total_steps = 0

    # reference no-ops (firstfit.py:9-19)
    def learn(self):
        pass

    def load_model(self, modelpath):
        pass

    def save_model(self, modelpath):
        pass

    def eval(self, model=True):
        pass

    def act(self, observation, out=None):
        """`out`: optional [n, V] tensor the kernel writes the actions into — a device tensor, or a PINNED host tensor
        (device-mapped under UVA: the kernel stores over PCIe directly, no copy-engine transfer)."""
        vec = self.vec
        single = False
        if isinstance(observation, np.ndarray):
            single = observation.ndim == 1
            obs = torch.from_numpy(np.ascontiguousarray(observation, dtype=np.float32).reshape(-1, vec.obs_dim)).to(vec.device)
        else:
            obs = observation.to(vec.device, torch.float32)
            single = obs.dim() == 1
            obs = obs.reshape(-1, vec.obs_dim).contiguous()
        n = obs.shape[0]
        host_out = isinstance(observation, np.ndarray)
        dtype = torch.int64 if host_out else vec.place_dtype
        if out is not None:
            if out.shape != (n, vec.V) or not out.is_contiguous() or not (out.is_cuda or out.is_pinned()):
                raise ValueError("out must be a contiguous [n, V] device or pinned-host tensor")
            dtype, host_out = out.dtype, False
        action = out if out is not None else torch.empty((n, vec.V), dtype=dtype, device=vec.device)
        code = {torch.uint8: nv.U8, torch.int16: nv.I16, torch.int64: nv.I64}[dtype]
        with torch.cuda.device(vec.device):
            nv.check(nv.lib().vmgym_agent_act(C.byref(vec._ccfg()), self.kind, nv.TIE_IDS[self.tiebreak], obs.data_ptr(), n,
                                              action.data_ptr(), code, vec._stream()), "vmgym_agent_act")
        if host_out:
            a = action.cpu().numpy()
            return a[0] if single else a
        return action[0] if single else action

    def _fused_kind(self):
        return {nv.AGENT_FIRSTFIT: "firstfit", nv.AGENT_BESTFIT: "bestfit"}.get(self.kind)


class FirstFitAgent(HeuristicAgent):
    """src/agents/firstfit.py:21-38."""
    kind = nv.AGENT_FIRSTFIT
    name = "FirstFitAgent"


class BestFitAgent(HeuristicAgent):
    """src/agents/bestfit.py:21-40 (tie rule: DESIGN.md "best-fit tie-break")."""
    kind = nv.AGENT_BESTFIT
    name = "BestFitAgent"
