"""VecVmEnv — N independent reference-semantics VmEnv instances resident on one B200.

Mirrors the reference env surface (vmenv/envs/env.py:19-325) with a leading env axis:
    reset(seed) -> obs[N, 3V+2P] f32                       (env.py:180-226)
    step(action[N, V]) -> obs, reward[N] f64, terminated[N] bool, truncated[N] bool, info  (env.py:66-103)
    get_invalid_action_mask(masked) -> bool[N, V, A]       (env.py:45-53)
    eval(mode), seed(seed), close()                        (env.py:105-106,172-178,241-242)
plus the fused heuristic rollouts `agent_step(kind, n_steps)` (Base.test loop, src/agents/base.py:71-86).

All state lives in one device buffer of env records (include/vmgym.h: vmgym_layout); every method enqueues
hand-written sm_100a kernels from libvmgym.so on the current CUDA stream.  Returned tensors are views of
buffers owned by the env and are overwritten by the next call (unlike the reference, which returns fresh
arrays) — clone them if they must survive a step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nv
from .config import Config
from .trace import EnvStreams, cdf_brackets, poisson_cdf_table, sample_numpy_traces, size_code_range

import contextlib

_TORCH_ACTION_DTYPES = {torch.uint8: nv.U8, torch.int16: nv.I16, torch.int64: nv.I64}
_NULL_CTX = contextlib.nullcontext()


def _hist_stats(h):
    """mean / median / max of the list that has h[k] copies of k / 1000.0 (np.mean / np.median / np.max semantics)."""
    n = int(h.sum())
    if n == 0:
        return 0.0, 0.0, 0.0
    ks = np.nonzero(h)[0]
    vals = ks / 1000.0
    mean = float(np.dot(h[ks], vals) / n)
    cum = np.cumsum(h)
    lo = int(np.searchsorted(cum, (n - 1) // 2 + 1))        # value at sorted index (n-1)//2
    hi = int(np.searchsorted(cum, n // 2 + 1))              # value at sorted index n//2
    median = (lo / 1000.0 + hi / 1000.0) / 2.0 if n % 2 == 0 else hi / 1000.0
    return mean, float(median), float(ks[-1] / 1000.0)


def vm_stats_from_histograms(hist, totals):
    """hist int64 [N, 2, BINS] (pending, slowdown), totals int64 [N, 4] (VMs, placed VMs, sum of lifetimes) ->
    Record.get_summary's per-VM keys (record.py:118-125) per env, unrounded."""
    N = hist.shape[0]
    out = {k: np.zeros(N) for k in ("average VM life", "average pending", "median pending", "max pending",
                                     "average slowdown", "median slowdown", "max slowdown", "vms", "placed vms")}
    for i in range(N):
        n_vms, n_placed, life = int(totals[i, 0]), int(totals[i, 1]), int(totals[i, 2])
        out["vms"][i], out["placed vms"][i] = n_vms, n_placed
        out["average VM life"][i] = life / n_vms if n_vms else float("nan")        # np.mean([]) is nan in the reference too
        out["average pending"][i], out["median pending"][i], out["max pending"][i] = _hist_stats(hist[i, 0])
        # record.py:83-84: no placed VM -> the slowdown list is [0]
        out["average slowdown"][i], out["median slowdown"][i], out["max slowdown"][i] = _hist_stats(hist[i, 1])
    return out


def ctypes_addr(obj) -> int:
    return C.addressof(obj)


class VecVmEnv:
    def __init__(self, config: Config, num_envs: int, device="cuda", rng: str = "numpy", seeds=None,
                 trace_steps: int | None = None, max_admissions: int | None = None, tiebreak: str = "stable",
                 env_configs=None):
        """`env_configs`: optional list of num_envs Configs that may differ from `config` in arrival_rate,
        service_length, sequence and seed only (rng="numpy"): each env draws its traces from its own parameters, so a
        sweep over load / service length / VM size mix / seed is ONE batch (exp_suspension.py:75-85,
        exp_vm_size.py:13-20, exp_performance.py:26,37 run one process per point)."""
        if not torch.cuda.is_available():
            raise nv.VmgymError("VecVmEnv needs a CUDA device (sm_100a); there is no CPU path")
        self.config = config.validate()
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise nv.VmgymError("VecVmEnv needs a CUDA device")
        self.rng_mode = rng
        if rng not in ("numpy", "philox"):
            raise ValueError("rng must be 'numpy' (reference-exact pre-sampled traces) or 'philox'")
        self.env_configs = None
        if env_configs is not None:
            if rng != "numpy" or len(env_configs) != int(num_envs):
                raise ValueError("env_configs needs rng='numpy' and one Config per env")
            same = ("pms", "vms", "training_steps", "eval_steps", "reward_function", "cap_target_util", "beta", "allow_null_action")
            for c in env_configs:
                if any(getattr(c, k) != getattr(config, k) for k in same):
                    raise ValueError(f"env_configs may only differ in arrival_rate / service_length / sequence / seed (not {same})")
            self.env_configs = [c.validate() for c in env_configs]
            if seeds is None:
                seeds = [int(c.seed) for c in env_configs]
        self.tiebreak = tiebreak
        self.eval_mode = False
        self.P, self.V = int(config.pms), int(config.vms)
        self.action_dim = config.action_dim
        self.obs_dim = config.obs_dim
        self.WAIT_STATUS, self.NULL_STATUS = self.P, self.P + 1
        self._lib = nv.lib()
        self._ccfg_cache = {}
        self._out_cache = {}
        self._vm_slots = self._vm_hist = self._vm_totals = None      # per-VM statistics buffers (enable_vm_stats)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self._dev_index)
        self._layout = nv.Layout()
        nv.check(self._lib.vmgym_get_layout(C.byref(self._ccfg()), C.byref(self._layout)), "vmgym_get_layout")
        L = self._layout
        self.place_dtype = torch.uint8 if L.place_bytes == 1 else torch.int16
        N = self.num_envs
        with self._on_device():
            self.state = torch.zeros((N, L.record_bytes), dtype=torch.uint8, device=self.device)
            self.obs = torch.zeros((N, self.obs_dim), dtype=torch.float32, device=self.device)
            self.reward = torch.zeros(N, dtype=torch.float64, device=self.device)
            self.terminated_u8 = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self.terminated = self.terminated_u8.view(torch.bool)      # same storage, reference dtype
            self.truncated = torch.zeros(N, dtype=torch.bool, device=self.device)   # always False (env.py:102)
            self.valid = torch.zeros((N, self.V), dtype=torch.uint8, device=self.device)
            self.agent_action = torch.zeros((N, self.V), dtype=self.place_dtype, device=self.device)
            self.stats = torch.zeros((N, nv.STATS), dtype=torch.float64, device=self.device)
        self._trace_steps = trace_steps
        self._max_admissions = max_admissions
        self._streams = None
        self._trace = None          # nv.Trace
        self._trace_tensors = ()
        self._seeds = None
        self._build_views()
        base = int(config.seed)
        self.reset(seed=(base + np.arange(N, dtype=np.int64)) if seeds is None else np.asarray(seeds, np.int64))

    # ------------------------------------------------------------------------------------------------
    def _ccfg(self) -> nv.Config:
        cached = self._ccfg_cache.get(self.eval_mode)
        if cached is None:
            c = self.config
            limit = int(c.eval_steps) if self.eval_mode else int(c.training_steps)
            cached = nv.Config(int(c.pms), int(c.vms), int(bool(c.allow_null_action)), nv.REWARD_IDS[c.reward_function],
                               int(bool(c.cap_target_util)), limit, float(c.beta))
            self._ccfg_cache[self.eval_mode] = cached
        return cached

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _on_device(self):
        """Context that makes self.device current (no-op when it already is: the common one-process-per-GPU case)."""
        if torch.cuda.current_device() == self._dev_index:
            return _NULL_CTX
        return torch.cuda.device(self.device)

    def _build_views(self):
        L, P, V = self._layout, self.P, self.V
        s = self.state
        self.cpu = s[:, L.off_cpu:L.off_cpu + 8 * P].view(torch.float64)
        self.memory = s[:, L.off_memory:L.off_memory + 8 * P].view(torch.float64)
        self._vm_remaining_u16 = s[:, L.off_remaining:L.off_remaining + 2 * V].view(torch.int16)   # raw u16 bits
        pb = L.place_bytes
        self.vm_placement = s[:, L.off_placement:L.off_placement + pb * V].view(self.place_dtype)
        self.vm_cpu_code = s[:, L.off_cpu_code:L.off_cpu_code + V]
        self.vm_mem_code = s[:, L.off_mem_code:L.off_mem_code + V]
        sc = s[:, L.off_scalars:L.off_scalars + nv.SCALARS_BYTES]
        self._scalars_i32 = sc[:, :40].view(torch.int32)
        self._scalars_i64 = sc[:, 40:80].view(torch.int64)     # seed, cpu_code_sum, mem_code_sum, (f64 x2 as bits)
        self._scalars_f64 = sc[:, 64:80].view(torch.float64)   # episode_return, last_reward

    @property
    def vm_remaining_runtime(self):
        """int32 [N, V]: vm_remaining_runtime of env.py:192.  The record keeps steps-left for waiting slots and the departure
        step modulo 2^16 for running slots (include/vmgym.h, off_remaining); this converts the latter back."""
        raw = self._vm_remaining_u16.to(torch.int32) & 0xffff
        running = (self.vm_placement.to(torch.int32) & 0xffff) < self.P
        t_next = self._scalars_i32[:, :1]
        return torch.where(running, ((raw - t_next) & 0xffff) + 1, raw)

    def vm_arrival_step(self):
        """int64 [N, V]: the step at which the VM occupying each slot arrived, as the reference logs it
        (`vm_arrival_steps[i].append(timestep + 1)`, env.py:293); 0 for slots never occupied.  Needs enable_vm_stats()."""
        if self._vm_slots is None:
            raise RuntimeError("call enable_vm_stats() (before reset) first")
        a = self._vm_slots[:, :, 0].to(torch.int64)
        return torch.where(a > 0, a + 1, a)

    # ------------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------------
    def eval(self, eval_mode: bool = True):
        self.eval_mode = bool(eval_mode)
        # pre-sampled traces cover the step limit of the mode they were drawn in (reset): a longer limit needs a reset
        if self.rng_mode == "numpy" and self._trace is not None and self._trace_steps is None:
            limit = int(self.config.eval_steps) if self.eval_mode else int(self.config.training_steps)
            self._trace_short = limit > int(self._trace.arrivals_len)

    def _check_trace(self):
        if getattr(self, "_trace_short", False):
            raise nv.VmgymError("eval() raised the step limit beyond the pre-sampled trace: call reset() (as Base.test does, "
                                "base.py:65-67), or construct with trace_steps=..., or use rng='philox'")

    def seed(self, seed=None):
        """env.py:172-178 — (re)create the four generators of every env at seed_i .. seed_i+3."""
        if seed is None:
            seed = int(self.config.seed) + np.arange(self.num_envs, dtype=np.int64)
        seeds = np.broadcast_to(np.asarray(seed, np.int64), (self.num_envs,)).copy() if np.ndim(seed) else \
            int(seed) + np.arange(self.num_envs, dtype=np.int64)
        self._seeds = seeds
        if self.rng_mode == "numpy":
            self._streams = [EnvStreams(int(s)) for s in seeds]
        self._reseeded = True

    def reset(self, seed=None, options=None):
        """env.py:180-226.  `seed` None continues the arrival/service streams; an int seeds env i with seed+i;
        an array gives one seed per env."""
        if seed is not None:
            self.seed(seed)
        rewind = 1
        d_seeds = None
        with self._on_device():
            if self.rng_mode == "numpy":
                if not getattr(self, "_reseeded", False) and self._trace is not None:
                    pos = self._scalars_i32[:, 6:8].cpu().numpy().astype(np.int64)
                    for i, (s, (a, j)) in enumerate(zip(self._streams, pos)):
                        s.rewind_to(self.env_configs[i] if self.env_configs else self.config, a, j)
                # one arrival draw per step of the episode: the active mode's step limit (the size sequences keep the
                # reference's 2 * max(training_steps, eval_steps) draws per generator, env.py:210-219)
                T = self._trace_steps or (int(self.config.eval_steps) if self.eval_mode else int(self.config.training_steps))
                self._trace_short = False
                arr, adm = sample_numpy_traces(self.env_configs or self.config, self._streams, T, self._max_admissions)
                d_arr = torch.from_numpy(arr.view(np.int16)).to(self.device)
                d_adm = torch.from_numpy(adm.view(np.int32)).to(self.device)
                self._trace_tensors = (d_arr, d_adm)
                self._trace = nv.Trace(mode=nv.TRACE_PRESAMPLED, d_arrivals=d_arr.data_ptr(), arrivals_len=arr.shape[1],
                                       d_admissions=d_adm.data_ptr(), admissions_len=adm.shape[1])
            else:
                if self._trace is None:
                    ka, ta = poisson_cdf_table(self.config.arrival_rate)
                    ks, ts = poisson_cdf_table(self.config.service_length)
                    d_ta = torch.from_numpy(ta.view(np.int64)).to(self.device)
                    d_ts = torch.from_numpy(ts.view(np.int64)).to(self.device)
                    lo, hi = size_code_range(self.config.sequence)
                    d_br = torch.from_numpy(cdf_brackets(ts).view(np.int16)).to(self.device) if len(ts) < 65536 else None
                    self._trace_tensors = (d_ta, d_ts, d_br)
                    self._trace = nv.Trace(mode=nv.TRACE_PHILOX, d_arrival_cdf=d_ta.data_ptr(), arrival_cdf_len=len(ta),
                                           arrival_kmin=ka, d_service_cdf=d_ts.data_ptr(), service_cdf_len=len(ts),
                                           service_kmin=ks, size_lo_code=lo, size_hi_code=hi,
                                           d_service_bracket=d_br.data_ptr() if d_br is not None else None)
                    self.philox_tables = (ka, ta, ks, ts, lo, hi)
                if getattr(self, "_reseeded", False):
                    d_seeds = torch.from_numpy(self._seeds.copy()).to(self.device)
                else:
                    rewind = 0
            self._reseeded = False
            self.stats.zero_()
            if self._vm_slots is not None:
                self._vm_slots.zero_(); self._vm_hist.zero_(); self._vm_totals.zero_()
            nv.check(self._lib.vmgym_reset(C.byref(self._ccfg()), self.state.data_ptr(), self.num_envs, None,
                                           d_seeds.data_ptr() if d_seeds is not None else None, rewind,
                                           self.obs.data_ptr(), self._stream()), "vmgym_reset")
        return self.obs, {}

    def _outputs(self, want_action=False, want_stats=False, want_obs=True, want_valid=True, lo: int = 0) -> nv.Outputs:
        """Output pointers of a launch over the envs [lo, ...): every per-env buffer offset to row `lo`."""
        key = (want_action, want_stats, want_obs, want_valid, lo)
        out = self._out_cache.get(key)
        if out is None:
            vm = self._vm_slots is not None
            out = nv.Outputs(d_obs=self.obs[lo:].data_ptr() if want_obs else None, d_reward=self.reward[lo:].data_ptr(),
                             d_terminated=self.terminated_u8[lo:].data_ptr(),
                             d_valid=self.valid[lo:].data_ptr() if want_valid else None,
                             d_action=self.agent_action[lo:].data_ptr() if want_action else None,
                             d_stats=self.stats[lo:].data_ptr() if want_stats else None,
                             d_vm_slots=self._vm_slots[lo:].data_ptr() if vm else None,
                             d_vm_hist=self._vm_hist[lo:].data_ptr() if vm else None,
                             d_vm_totals=self._vm_totals[lo:].data_ptr() if vm else None,
                             obs_persistent=1)          # self.obs is this env's own buffer: unchanged envs keep their rows
            self._out_cache[key] = out
        return out

    def _trace_at(self, lo: int):
        """The trace descriptor for a launch over the envs [lo, ...): pre-sampled rows are per env, Philox tables are shared."""
        if lo == 0 or self.rng_mode != "numpy":
            return self._trace
        t = nv.Trace.from_buffer_copy(self._trace)
        d_arr, d_adm = self._trace_tensors
        t.d_arrivals, t.d_admissions = d_arr[lo:].data_ptr(), d_adm[lo:].data_ptr()
        return t

    def step(self, action, want_obs: bool = True, want_valid: bool = True, host_outputs=None, want_stats: bool = False,
             obs_mirror=None, next_action=None):
        """env.py:66-103 for all envs.  `action`: [N, V] device tensor (uint8 / int16 / int64), numpy int array, or a
        PINNED host tensor (device-mapped under UVA: the kernel reads it over PCIe, no copy-engine transfer).
        `host_outputs`: optional (reward f64 [N], terminated u8 [N]) pinned host tensors the kernel writes directly,
        instead of the env's device buffers.  `want_stats`: accumulate the episode sums behind summary().
        `obs_mirror`: optional pinned host tensor [N, 3V+2P] kept equal to self.obs by storing only the entries that changed
        (it must already equal self.obs, e.g. copied once after reset).
        `next_action`: optional (agent name, [N, V] device tensor of the placement dtype, tiebreak or None): the kernel also
        writes that heuristic agent's act() on the NEW state into the tensor (what `agent.act(obs)` of the returned observation
        gives; the tensor may be `action` itself when the dtypes agree)."""
        if not isinstance(action, torch.Tensor):
            action = torch.from_numpy(np.ascontiguousarray(action, dtype=np.int64)).to(self.device, non_blocking=True)
        elif not action.is_cuda and not action.is_pinned():
            action = action.to(self.device, non_blocking=True)
        if action.shape != (self.num_envs, self.V):
            raise ValueError(f"action must have shape {(self.num_envs, self.V)}, got {tuple(action.shape)}")
        if action.dtype not in _TORCH_ACTION_DTYPES:
            action = action.to(torch.int64)
        action = action.contiguous()
        self._check_trace()
        out = self._outputs(want_obs=want_obs, want_valid=want_valid, want_stats=want_stats)
        if host_outputs is not None:
            rew_h, term_h = host_outputs
            if not (rew_h.is_pinned() and term_h.is_pinned() and rew_h.dtype == torch.float64 and term_h.dtype == torch.uint8
                    and rew_h.numel() == self.num_envs and term_h.numel() == self.num_envs):
                raise ValueError("host_outputs must be pinned (float64 [N], uint8 [N]) tensors")
            if obs_mirror is not None and not (obs_mirror.is_pinned() and obs_mirror.dtype == torch.float32 and want_obs
                                               and obs_mirror.shape == (self.num_envs, self.obs_dim) and obs_mirror.is_contiguous()):
                raise ValueError("obs_mirror must be a contiguous pinned float32 [N, 3V+2P] tensor (and want_obs)")
            key = ("host", rew_h.data_ptr(), term_h.data_ptr(), want_obs, want_valid, obs_mirror.data_ptr() if obs_mirror is not None else 0)
            hout = self._out_cache.get(key)
            if hout is None:
                hout = nv.Outputs(d_obs=out.d_obs, d_reward=rew_h.data_ptr(), d_terminated=term_h.data_ptr(), d_valid=out.d_valid,
                                  d_action=None, d_stats=None, d_vm_slots=out.d_vm_slots, d_vm_hist=out.d_vm_hist,
                                  d_vm_totals=out.d_vm_totals, d_obs_mirror=obs_mirror.data_ptr() if obs_mirror is not None else None,
                                  obs_persistent=1)
                self._out_cache[key] = hout
            out = hout
        if next_action is not None:
            na_agent, na_out, na_tie = next_action
            if not (na_out.is_cuda and na_out.dtype == self.place_dtype and na_out.shape == (self.num_envs, self.V) and na_out.is_contiguous()):
                raise ValueError("next_action: need a contiguous device tensor [N, V] of the placement dtype")
            key = ("next", ctypes_addr(out), na_out.data_ptr(), na_agent, na_tie)
            nout = self._out_cache.get(key)
            if nout is None:
                nout = nv.Outputs.from_buffer_copy(out)
                nout.d_next_action = na_out.data_ptr()
                nout.next_agent = {"firstfit": nv.AGENT_FIRSTFIT, "bestfit": nv.AGENT_BESTFIT}[na_agent]
                nout.next_tiebreak = nv.TIE_IDS[na_tie or self.tiebreak]
                self._out_cache[key] = nout
            out = nout
        with self._on_device():
            nv.check(self._lib.vmgym_step(C.byref(self._ccfg()), self.state.data_ptr(), self.num_envs,
                                          C.byref(self._trace), action.data_ptr(), _TORCH_ACTION_DTYPES[action.dtype],
                                          C.byref(out), self._stream()), "vmgym_step")
        return self.obs, self.reward, self.terminated, self.truncated, {"action": action, "valid": self.valid}

    def agent_step(self, agent: str, n_steps: int = 1, want_obs: bool = True, want_action: bool = True,
                   want_stats: bool = False, want_valid: bool = True, tiebreak: str | None = None, envs: tuple | None = None):
        """Fused heuristic agent.act + env.step, `n_steps` per launch (firstfit.py:21-38 / bestfit.py:21-40 +
        env.py:66-103).  `envs=(lo, hi)` restricts the launch to that contiguous range of envs (the others are untouched).
        Returns (obs, reward, terminated) of the last executed step."""
        kind = {"firstfit": nv.AGENT_FIRSTFIT, "bestfit": nv.AGENT_BESTFIT}[agent]
        self._check_trace()
        lo, hi = (0, self.num_envs) if envs is None else (int(envs[0]), int(envs[1]))
        if not 0 <= lo <= hi <= self.num_envs:
            raise ValueError("envs must be a range inside [0, num_envs]")
        out = self._outputs(want_action=want_action, want_stats=want_stats, want_obs=want_obs, want_valid=want_valid, lo=lo)
        trace = self._trace_at(lo)
        with self._on_device():
            nv.check(self._lib.vmgym_agent_step(C.byref(self._ccfg()), self.state[lo:].data_ptr(), hi - lo,
                                                C.byref(trace), kind, nv.TIE_IDS[tiebreak or self.tiebreak],
                                                int(n_steps), C.byref(out), self._stream()), "vmgym_agent_step")
        return self.obs, self.reward, self.terminated

    def agent_step_rotation(self, agent: str, batch_envs: int, batch_steps: int, first_batch: int = 0, n_steps: int = 1,
                            want_obs: bool = True, want_action: bool = False, want_stats: bool = False, want_valid: bool = False,
                            tiebreak: str | None = None):
        """The env batch seen as num_envs / batch_envs independent sub-batches (e.g. the seeds or load points of a sweep,
        exp_performance.py:63-83): ONE persistent launch runs `batch_steps` consecutive batch steps, batch step k advancing
        sub-batch (first_batch + k) % n_batches by `n_steps` fused act + env.step.  Same results as calling agent_step on
        the sub-batches in rotation, without a launch boundary per step.  Returns the index of the next sub-batch in turn."""
        kind = {"firstfit": nv.AGENT_FIRSTFIT, "bestfit": nv.AGENT_BESTFIT}[agent]
        if batch_envs < 1 or self.num_envs % batch_envs:
            raise ValueError("batch_envs must divide num_envs")
        nb = self.num_envs // batch_envs
        out = self._outputs(want_action=want_action, want_stats=want_stats, want_obs=want_obs, want_valid=want_valid)
        with self._on_device():
            nv.check(self._lib.vmgym_agent_step_rotation(C.byref(self._ccfg()), self.state.data_ptr(), int(batch_envs), nb,
                                                         int(first_batch) % nb, int(batch_steps), C.byref(self._trace), kind,
                                                         nv.TIE_IDS[tiebreak or self.tiebreak], int(n_steps), C.byref(out),
                                                         self._stream()), "vmgym_agent_step_rotation")
        return (int(first_batch) + int(batch_steps)) % nb

    def evaluate(self, agent: str, seeds=None, chunk: int = 1000, tiebreak: str | None = None):
        """Base.test (src/agents/base.py:63-124) for a fused heuristic agent, entirely on the device: eval mode,
        reset(seed), act/step until the episode terminates, with the running sums behind Record.get_summary
        (src/record.py:110-134) and the columns of the published tables (exp_performance.py:104-147) accumulated in the
        kernel.  Returns a dict of numpy arrays [N]."""
        self.eval(True)
        if seeds is None:
            seeds = [int(c.seed) for c in self.env_configs] if self.env_configs else self.config.seed + np.arange(self.num_envs)
        self.reset(seed=np.asarray(seeds, np.int64))
        limit = int(self.config.eval_steps)
        done = 0
        while done < limit:
            n = min(chunk, limit - done)
            self.agent_step(agent, n_steps=n, want_obs=False, want_action=False, want_valid=False, want_stats=True,
                            tiebreak=tiebreak)
            done += n
        return self.summary()

    def summary(self):
        """Record.get_summary (src/record.py:110-134) of the episode so far, per env (unrounded float64 / int64 numpy arrays
        [N]); needs the steps to have run with want_stats (evaluate does) and, for the per-VM keys, enable_vm_stats()."""
        c = self.counters()
        st = self.stats.cpu().numpy()
        steps = np.maximum(st[:, 7], 1.0)
        per_vm = self.vm_stats() if self._vm_slots is not None else {}
        return {**per_vm, "total rewards": c["episode_return"], "total served VMs": c["served_requests"], "total requests": c["total_requests"],
                "total cpu requested": c["total_cpu_requested"], "total memory requested": c["total_memory_requested"],
                "total suspend actions": c["suspend_actions"], "total place actions": c["place_actions"],
                "dropped requests": c["dropped_requests"], "drop rate": st[:, 0] / steps, "waiting ratio": st[:, 1] / steps,
                "cpu mean": st[:, 2] / steps, "cpu var": st[:, 3] / steps, "memory mean": st[:, 4] / steps,
                "memory var": st[:, 5] / steps, "rejected actions": st[:, 6], "steps": st[:, 7],
                # np.std over the whole T x P matrix (record.py:128,131): E[x^2] - mean^2 with E[x^2] = mean_t(var_p + mean_p^2)
                "cpu std": np.sqrt(np.maximum((st[:, 3] + st[:, 8]) / steps - (st[:, 2] / steps) ** 2, 0.0)),
                "memory std": np.sqrt(np.maximum((st[:, 5] + st[:, 9]) / steps - (st[:, 4] / steps) ** 2, 0.0)),
                "cpu mean target": st[:, 10] / steps, "memory mean target": st[:, 11] / steps, "rank mean": st[:, 12] / steps}

    # ---- Record's per-VM statistics (src/record.py:34-96,110-134) -------------------------------------------
    def enable_vm_stats(self):
        """Track, for every VM that occupies a slot, what Record derives from the V x T placement samples: pending rate,
        slowdown rate and lifetime.  Takes effect from the next reset() (buffers are zeroed there); the step kernels keep
        four clocks per slot and bin a VM's rates when it departs.  Steps then run through the generic kernel."""
        if self._vm_slots is None:
            N, V = self.num_envs, self.V
            self._vm_slots = torch.zeros((N, V, 4), dtype=torch.int32, device=self.device)
            self._vm_hist = torch.zeros((N, 2, nv.VMSTAT_BINS), dtype=torch.int32, device=self.device)
            self._vm_totals = torch.zeros((N, 4), dtype=torch.int64, device=self.device)
            self._out_cache.clear()
        return self

    def vm_stats(self):
        """The per-VM keys of Record.get_summary for the episode so far, per env (float64 numpy arrays [N]): VMs that
        departed plus the ones still in a slot (their sample list ends at the last step, as in Record)."""
        if self._vm_slots is None:
            raise RuntimeError("call enable_vm_stats() (before reset) first")
        N = self.num_envs
        hist = torch.empty_like(self._vm_hist)
        totals = torch.empty_like(self._vm_totals)
        with self._on_device():
            nv.check(self._lib.vmgym_vmstats_finalize(C.byref(self._ccfg()), self.state.data_ptr(), N, self._vm_slots.data_ptr(),
                                                      self._vm_hist.data_ptr(), self._vm_totals.data_ptr(), hist.data_ptr(),
                                                      totals.data_ptr(), self._stream()), "vmgym_vmstats_finalize")
        h = hist.cpu().numpy().astype(np.int64)
        t = totals.cpu().numpy()
        return vm_stats_from_histograms(h, t)

    def capture(self, fn, warmup: int = 2):
        """Capture `fn` (a callable issuing env calls with fixed arguments, e.g. one fused agent step) into a CUDA
        graph; returns the torch.cuda.CUDAGraph — `.replay()` re-issues the launches without host-side overhead."""
        with self._on_device():
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    fn()
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fn()
        return graph

    def observe(self):
        with self._on_device():
            nv.check(self._lib.vmgym_observe(C.byref(self._ccfg()), self.state.data_ptr(), self.num_envs,
                                             self.obs.data_ptr(), self._stream()), "vmgym_observe")
        return self.obs

    def get_invalid_action_mask(self, masked: bool = True):
        """env.py:45-53: bool[N, V, A], True = invalid (all False when `masked` is False)."""
        mask = torch.zeros((self.num_envs, self.V, self.action_dim), dtype=torch.uint8, device=self.device)
        if masked:
            with self._on_device():
                nv.check(self._lib.vmgym_invalid_action_mask(C.byref(self._ccfg()), self.state.data_ptr(), self.num_envs,
                                                             mask.data_ptr(), self._stream()), "vmgym_invalid_action_mask")
        return mask.bool()

    def close(self):
        pass

    # ------------------------------------------------------------------------------------------------
    # state access (device views / host snapshots)
    # ------------------------------------------------------------------------------------------------
    def counters(self):
        """dict of int64 numpy arrays [N]: the scalar attributes of env.py:197-208."""
        i32 = self._scalars_i32.cpu().numpy().astype(np.int64)
        i64 = self._scalars_i64.cpu().numpy()
        f64 = self._scalars_f64.cpu().numpy()
        d = {k: i32[:, i] for i, k in enumerate(nv.SCALARS_I32)}
        d["seed"] = i64[:, 0]
        d["total_cpu_requested"] = i64[:, 1] / 100.0
        d["total_memory_requested"] = i64[:, 2] / 100.0
        d["episode_return"] = f64[:, 0]
        d["last_reward"] = f64[:, 1]
        return d

    def state_dict_host(self, env: int = 0):
        """Host snapshot of one env in the reference's attribute names/dtypes (for parity tests / the N=1 facade)."""
        P, V = self.P, self.V
        rec = self.state[env].cpu().numpy()
        L = self._layout
        place = rec[L.off_placement:L.off_placement + L.place_bytes * V].view(np.uint8 if L.place_bytes == 1 else np.uint16)
        cc = rec[L.off_cpu_code:L.off_cpu_code + V]
        mc = rec[L.off_mem_code:L.off_mem_code + V]
        i32 = rec[L.off_scalars:L.off_scalars + 40].view(np.int32)
        i64 = rec[L.off_scalars + 40:L.off_scalars + 64].view(np.int64)
        f64 = rec[L.off_scalars + 64:L.off_scalars + 80].view(np.float64)
        placement = place.astype(np.int64)
        vm_cpu = (cc & 0x7f) / 100.0
        vm_memory = mc / 100.0
        existing = placement <= P
        n_ex = int(existing.sum())
        s = dict(vm_placement=placement, vm_cpu=vm_cpu, vm_memory=vm_memory,
                 cpu=rec[L.off_cpu:L.off_cpu + 8 * P].view(np.float64).copy(),
                 memory=rec[L.off_memory:L.off_memory + 8 * P].view(np.float64).copy(),
                 vm_remaining_runtime=np.where(placement < P, ((rec[L.off_remaining:L.off_remaining + 2 * V].view(np.uint16).astype(np.int64)
                                                               - int(i32[0])) & 0xffff) + 1,
                                               rec[L.off_remaining:L.off_remaining + 2 * V].view(np.uint16).astype(np.int64)),
                 vm_suspended=(cc >> 7).astype(np.int64))
        for i, k in enumerate(nv.SCALARS_I32):
            s[k] = int(i32[i])
        s["suspend_action"], s["place_action"] = s["suspend_actions"], s["place_actions"]
        s["arr_cursor"], s["adm_cursor"], s["trace_exhausted"] = s["arrival_pos"], s["admission_pos"], s["status"] & 1
        s["total_cpu_requested"] = int(i64[1]) / 100.0
        s["total_memory_requested"] = int(i64[2]) / 100.0
        s["episode_return"], s["last_reward"] = float(f64[0]), float(f64[1])
        # derived metrics of env.py:112-121 (not stored on the device)
        s["waiting_ratio"] = float((placement == P).sum()) / n_ex if n_ex else 0
        tc = float(np.sum((cc & 0x7f)[existing].astype(np.int64))) / 100.0 / P
        tm = float(np.sum(mc[existing].astype(np.int64))) / 100.0 / P
        if self.config.cap_target_util:
            tc, tm = min(tc, 1.0), min(tm, 1.0)
        s["target_cpu_mean"], s["target_memory_mean"] = tc, tm
        return s
