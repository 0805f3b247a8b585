"""ctypes front-end of the CPU oracle (oracle/vmenv_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of vmenv_oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` legs; never by the product
package.  `OracleVmEnv` mirrors the reference `VmEnv` surface (vmenv/envs/env.py:19-325) closely enough
that the parity tests read like the reference's own eval loop (src/agents/base.py:63-86).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvmoracle.so")

REWARD_IDS = {"wr": 1, "ut": 2, "kl": 3}          # main.py:94 order (reward 1/2/3)
SEQUENCES = {"uniform": (0.1, 1.0), "lowuniform": (0.1, 0.65), "highuniform": (0.25, 1.0)}  # env.py:211-219
TIE_STABLE, TIE_NUMPY_INTROSORT = 0, 1
AGENT_NOOP, AGENT_FIRSTFIT, AGENT_BESTFIT = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "vmenv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i64p, f64p, f32p, i32p, u8p = (C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_float),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_uint8))
        L.vmo_create.restype = C.c_void_p
        L.vmo_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int64, C.c_int64]
        L.vmo_destroy.argtypes = [C.c_void_p]
        L.vmo_set_trace.argtypes = [C.c_void_p, i32p, C.c_int64, f64p, f64p, i64p, C.c_int64]
        L.vmo_set_eval.argtypes = [C.c_void_p, C.c_int]
        L.vmo_reset.argtypes = [C.c_void_p]
        L.vmo_set_record.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.vmo_record_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.vmo_invalid_action_mask.argtypes = [C.c_void_p, C.c_int, u8p]
        L.vmo_get_obs.argtypes = [C.c_void_p, f32p]
        L.vmo_step.restype = C.c_int
        L.vmo_step.argtypes = [C.c_void_p, i64p, i64p, f32p, f64p]
        L.vmo_firstfit_act.argtypes = [C.c_int, C.c_int, f32p, i64p, f32p]
        L.vmo_bestfit_act.argtypes = [C.c_int, C.c_int, f32p, i64p, C.c_int, f32p, i64p]
        L.vmo_argsort_introsort_f32.argtypes = [f32p, i64p, C.c_int64]
        L.vmo_rollout.restype = C.c_int64
        L.vmo_rollout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, f64p]
        L.vmo_get_state.argtypes = [C.c_void_p, i64p, f64p, f64p, f64p, f64p, i64p, i64p, i64p, f64p]
        u64p = C.POINTER(C.c_uint64)
        L.vmo_philox_trace.argtypes = [C.c_uint64, C.c_int64, C.c_int64, u64p, C.c_int, C.c_int, u64p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, i32p, f64p, f64p, i64p]
        L.vmo_np_sum.restype = C.c_double
        L.vmo_np_sum.argtypes = [f64p, C.c_int64]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@dataclass
class OracleConfig:
    """Same 12 fields and defaults as the reference dataclass (vmenv/envs/config.py:3-16)."""
    arrival_rate: float = 0.182
    service_length: float = 100
    pms: int = 10
    vms: int = 30
    training_steps: int = 500
    eval_steps: int = 100000
    seed: int = 0
    reward_function: str = "wr"
    sequence: str = "uniform"
    cap_target_util: bool = True
    beta: float = 0.5
    allow_null_action: bool = False


class Trace:
    """Pre-sampled randomness of one episode, drawn exactly like the reference draws it
    (env.py:172-178 generators seed..seed+3; :211-219 size sequences of length 2*max_steps rounded to 2
    decimals; :272 one Poisson(arrival_rate) per step; :289 Poisson(service_length)+1 per admission)."""

    def __init__(self, arrivals, cpu_seq, mem_seq, svc_seq):
        self.arrivals = np.ascontiguousarray(arrivals, dtype=np.int32)
        self.cpu_seq = np.ascontiguousarray(cpu_seq, dtype=np.float64)
        self.mem_seq = np.ascontiguousarray(mem_seq, dtype=np.float64)
        self.svc_seq = np.ascontiguousarray(svc_seq, dtype=np.int64)


class _Streams:
    """The four numpy generators of one env (env.py:175-178) with replayable consumption, so that
    `reset()` without a seed continues the arrival/service streams where the episode left them."""

    def __init__(self, seed):
        self.rng = [np.random.default_rng(int(seed) + i) for i in range(4)]

    def sample(self, cfg, n_steps=None, n_adm=None) -> Trace:
        lo, hi = SEQUENCES[cfg.sequence]
        max_steps = max(int(cfg.training_steps), int(cfg.eval_steps))
        cpu_seq = np.around(self.rng[0].uniform(low=lo, high=hi, size=max_steps * 2), decimals=2)
        mem_seq = np.around(self.rng[1].uniform(low=lo, high=hi, size=max_steps * 2), decimals=2)
        T = max_steps if n_steps is None else int(n_steps)
        J = max_steps * 2 if n_adm is None else int(n_adm)
        self._state3 = self.rng[2].bit_generator.state
        self._state4 = self.rng[3].bit_generator.state
        arrivals = self.rng[2].poisson(cfg.arrival_rate, size=T)
        svc = self.rng[3].poisson(cfg.service_length, size=J) + 1
        return Trace(arrivals, cpu_seq[:J] if J <= cpu_seq.size else cpu_seq, mem_seq[:J] if J <= mem_seq.size else mem_seq, svc)

    def rewind_to(self, cfg, steps_used, adm_used):
        """Leave rng3/rng4 exactly where the reference would have left them after the episode."""
        self.rng[2].bit_generator.state = self._state3
        self.rng[3].bit_generator.state = self._state4
        if steps_used:
            self.rng[2].poisson(cfg.arrival_rate, size=int(steps_used))
        if adm_used:
            self.rng[3].poisson(cfg.service_length, size=int(adm_used))


def sample_trace(cfg, seed, n_steps=None, n_adm=None) -> Trace:
    return _Streams(seed).sample(cfg, n_steps, n_adm)


class OracleVmEnv:
    """Reference-shaped single env backed by the C restatement."""

    def __init__(self, config, trace_steps=None, trace_adm=None):
        self.config = config
        self.eval_mode = False
        P, V = int(config.pms), int(config.vms)
        self.P, self.V = P, V
        self.action_dim = P + 2 if config.allow_null_action else P + 1
        self.WAIT_STATUS, self.NULL_STATUS = P, P + 1
        self.obs_dim = 3 * V + 2 * P
        self._trace_steps, self._trace_adm = trace_steps, trace_adm
        self._h = lib().vmo_create(P, V, int(bool(config.allow_null_action)), REWARD_IDS[config.reward_function],
                                   float(config.beta), int(bool(config.cap_target_util)), int(config.training_steps),
                                   int(config.eval_steps))
        self._streams = None
        self._trace = None
        self.reset(config.seed)

    def __del__(self):
        try:
            lib().vmo_destroy(self._h)
        except Exception:
            pass

    # -- reference API ---------------------------------------------------------------------------
    def seed(self, seed=None):
        self._streams = _Streams(self.config.seed if seed is None else seed)

    def eval(self, eval_mode=True):
        self.eval_mode = bool(eval_mode)
        lib().vmo_set_eval(self._h, int(self.eval_mode))

    def set_trace(self, trace: Trace):
        self._trace = trace
        lib().vmo_set_trace(self._h, _p(trace.arrivals, C.c_int32), trace.arrivals.size, _p(trace.cpu_seq, C.c_double),
                            _p(trace.mem_seq, C.c_double), _p(trace.svc_seq, C.c_int64),
                            min(trace.cpu_seq.size, trace.mem_seq.size, trace.svc_seq.size))

    def reset(self, seed=None, options=None, trace: Trace | None = None):
        if trace is None:
            if seed is not None:
                self.seed(seed)
            elif self._trace is not None:
                c = self.counters()
                self._streams.rewind_to(self.config, c["arr_cursor"], c["adm_cursor"])
            trace = self._streams.sample(self.config, self._trace_steps, self._trace_adm)
        self.set_trace(trace)
        lib().vmo_reset(self._h)
        return self._obs(), {}

    def step(self, action):
        action = np.ascontiguousarray(action, dtype=np.int64)
        assert action.shape == (self.V,)
        valid = np.zeros(self.V, dtype=np.int64)
        obs = np.empty(self.obs_dim, dtype=np.float32)
        reward = C.c_double()
        term = lib().vmo_step(self._h, _p(action, C.c_int64), _p(valid, C.c_int64), _p(obs, C.c_float), C.byref(reward))
        return obs, reward.value, bool(term), False, {"action": action.copy(), "valid": valid}

    def get_invalid_action_mask(self, masked=True):
        m = np.zeros((self.V, self.action_dim), dtype=np.uint8)
        lib().vmo_invalid_action_mask(self._h, int(masked), _p(m, C.c_uint8))
        return m.astype(bool)

    def close(self):
        pass

    # -- state access ----------------------------------------------------------------------------
    def _obs(self):
        obs = np.empty(self.obs_dim, dtype=np.float32)
        lib().vmo_get_obs(self._h, _p(obs, C.c_float))
        return obs

    def state(self):
        V, P = self.V, self.P
        s = dict(vm_placement=np.empty(V, np.int64), vm_cpu=np.empty(V), vm_memory=np.empty(V), cpu=np.empty(P),
                 memory=np.empty(P), vm_remaining_runtime=np.empty(V, np.int64), vm_suspended=np.empty(V, np.int64))
        counters = np.empty(10, np.int64)
        scalars = np.empty(5, np.float64)
        lib().vmo_get_state(self._h, _p(s["vm_placement"], C.c_int64), _p(s["vm_cpu"], C.c_double),
                            _p(s["vm_memory"], C.c_double), _p(s["cpu"], C.c_double), _p(s["memory"], C.c_double),
                            _p(s["vm_remaining_runtime"], C.c_int64), _p(s["vm_suspended"], C.c_int64),
                            _p(counters, C.c_int64), _p(scalars, C.c_double))
        names = ["timestep", "total_requests", "served_requests", "suspend_action", "place_action", "dropped_requests",
                 "arr_cursor", "adm_cursor", "trace_exhausted"]
        s.update({k: int(counters[i]) for i, k in enumerate(names)})
        s.update(total_cpu_requested=scalars[0], total_memory_requested=scalars[1], waiting_ratio=scalars[2],
                 target_cpu_mean=scalars[3], target_memory_mean=scalars[4])
        return s

    def counters(self):
        s = self.state()
        return {k: v for k, v in s.items() if isinstance(v, int)}

    # -- Record (src/record.py) --------------------------------------------------------------------
    def enable_record(self, max_steps: int, max_arrivals: int | None = None):
        """Keep what Base.record_testing_step keeps for the per-VM statistics (base.py:135,140): the placement vector
        after every step and the env's vm_arrival_steps lists.  Rewound by reset()."""
        self._rec_log = np.zeros((int(max_steps), self.V), np.int16)
        self._arr_log = np.zeros((int(max_arrivals or 4 * max_steps + self.V), 2), np.int32)
        lib().vmo_set_record(self._h, self._rec_log.ctypes.data, self._rec_log.shape[0], self._arr_log.ctypes.data,
                             self._arr_log.shape[0])

    def record_lists(self):
        """Record.unique_vms_placement / pending_rates / slowdown_rates / vm_lifetime (record.py:34-96), restated
        line by line on the logged samples.  Returns (pending list, slowdown list, lifetime list)."""
        n_steps, n_arr = C.c_int64(), C.c_int64()
        lib().vmo_record_counts(self._h, C.byref(n_steps), C.byref(n_arr))
        assert n_steps.value <= self._rec_log.shape[0] and n_arr.value < self._arr_log.shape[0], "record log overflow"
        placements = np.transpose(self._rec_log[:n_steps.value].astype(np.int64))     # row is vm, col is timestep (:37)
        arrival_steps = [[] for _ in range(self.V)]
        for slot, step in self._arr_log[:n_arr.value]:
            arrival_steps[int(slot)].append(int(step))
        WAIT = self.WAIT_STATUS
        unique = []
        for vm, vm_status in enumerate(placements):                                   # :38-51
            if len(arrival_steps[vm]) == 0:
                continue
            start = 0
            for end in arrival_steps[vm][1:]:
                end -= 2                                                              # vm_placements starts at timestep 2
                spline = vm_status[start:end]
                unique.append(spline[spline <= WAIT])
                start = end
            spline = vm_status[start:]
            unique.append(spline[spline <= WAIT])
        pending, slowdown, life = [], [], []
        for status in unique:
            running = np.where(status < WAIT)[0]
            allocated_at = running[0] if running.size > 0 else None
            if allocated_at:                                                          # :60,75,91 (index 0 counts as None)
                pending.append(np.around((allocated_at + 1.0) / len(status), 3))
                slowdown_steps = np.count_nonzero(status[allocated_at:] == WAIT)
                vm_life = len(status) - allocated_at - 1
                slowdown.append(0 if vm_life == 0 else np.around(slowdown_steps / vm_life, 3))
                life.append(len(status) - allocated_at - 1)
            else:
                pending.append(1.0)
                life.append(0)
        if len(slowdown) == 0:
            slowdown = [0]                                                            # :83-84
        return pending, slowdown, life

    def record_summary(self):
        """The per-VM keys of Record.get_summary (record.py:118-125)."""
        pending, slowdown, life = self.record_lists()
        return {"average VM life": np.round(np.mean(life), 3), "average pending": np.round(np.mean(pending), 3),
                "median pending": np.round(np.median(pending), 3),
                "max pending": np.round(np.max(pending), 3) if len(pending) > 0 else 0,
                "average slowdown": np.round(np.mean(slowdown), 3), "median slowdown": np.round(np.median(slowdown), 3),
                "max slowdown": np.round(np.max(slowdown), 3)}

    def rollout(self, agent: int, steps: int, tiebreak: int = TIE_STABLE):
        """Base.test-style loop in C.  Returns (steps_run, stats dict)."""
        stats = np.zeros(16, np.float64)
        n = lib().vmo_rollout(self._h, int(agent), int(tiebreak), int(steps), _p(stats, C.c_double))
        keys = ["return", "served", "total_requests", "suspend", "place", "dropped", "drop_rate_mean", "cpu_mean",
                "cpu_var", "mem_mean", "mem_var", "waiting_ratio_mean", "steps", "rejected", "total_cpu_requested",
                "total_memory_requested"]
        return int(n), dict(zip(keys, stats.tolist()))


def firstfit_act(P, V, obs):
    """src/agents/firstfit.py:21-38 on a float32 observation."""
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    action = np.empty(V, np.int64)
    scratch = np.empty(P, np.float32)
    lib().vmo_firstfit_act(P, V, _p(obs, C.c_float), _p(action, C.c_int64), _p(scratch, C.c_float))
    return action


def bestfit_act(P, V, obs, tiebreak=TIE_STABLE):
    """src/agents/bestfit.py:21-40 on a float32 observation (tie rule: SURVEY §8c ruling)."""
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    action = np.empty(V, np.int64)
    scratch = np.empty(3 * P, np.float32)
    perm = np.empty(P, np.int64)
    lib().vmo_bestfit_act(P, V, _p(obs, C.c_float), _p(action, C.c_int64), int(tiebreak), _p(scratch, C.c_float),
                          _p(perm, C.c_int64))
    return action


def argsort_introsort_f32(keys):
    keys = np.ascontiguousarray(keys, dtype=np.float32)
    out = np.empty(keys.size, np.int64)
    lib().vmo_argsort_introsort_f32(_p(keys, C.c_float), _p(out, C.c_int64), keys.size)
    return out


def np_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().vmo_np_sum(_p(a, C.c_double), a.size)


def philox_trace(seed, n_steps, n_adm, arr_kmin, arr_cdf, svc_kmin, svc_cdf, lo_code, hi_code) -> Trace:
    """The draws the CUDA kernel makes in VMGYM_TRACE_PHILOX mode, as a pre-sampled Trace for the oracle env."""
    arr_cdf = np.ascontiguousarray(arr_cdf, np.uint64)
    svc_cdf = np.ascontiguousarray(svc_cdf, np.uint64)
    arrivals = np.empty(n_steps, np.int32)
    cpu, mem, svc = np.empty(n_adm), np.empty(n_adm), np.empty(n_adm, np.int64)
    lib().vmo_philox_trace(int(seed), n_steps, n_adm, _p(arr_cdf, C.c_uint64), arr_cdf.size, int(arr_kmin),
                           _p(svc_cdf, C.c_uint64), svc_cdf.size, int(svc_kmin), int(lo_code), int(hi_code),
                           _p(arrivals, C.c_int32), _p(cpu, C.c_double), _p(mem, C.c_double), _p(svc, C.c_int64))
    return Trace(arrivals, cpu, mem, svc)
