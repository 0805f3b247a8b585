"""Randomness of the env, host side.

`numpy` mode reproduces the reference's draws exactly (vmenv/envs/env.py:172-178: four PCG64 generators at
seed..seed+3; :211-219 size sequences rounded to 2 decimals; :272 one Poisson(arrival_rate) per step; :289
Poisson(service_length)+1 per admitted request) and packs them for the device:
    arrivals   u16[N, T]
    admissions u32[N, J] = cpu_code | mem_code << 8 | service << 16     (codes are hundredths)
Because admissions are consumed at a variable rate, the device keeps per-env cursors.

`philox` mode only builds the inverse-CDF tables the kernel samples from with its counter-based generator.
"""
from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .config import SEQUENCE_CODES, SEQUENCE_RANGES


class EnvStreams:
    """The four generators of one env, with replay so `reset()` without a seed continues rng3/rng4 exactly
    where the reference would have left them (drlvmp.py:450-452 calls seed() then reset())."""

    def __init__(self, seed: int):
        self.seed = int(seed)
        self.rng = [np.random.default_rng(self.seed + i) for i in range(4)]
        self._mark = None

    def draw(self, cfg, n_steps: int, n_adm: int | None):
        lo, hi = SEQUENCE_RANGES[cfg.sequence]
        max_steps = max(int(cfg.training_steps), int(cfg.eval_steps))
        self._mark = (self.rng[2].bit_generator.state, self.rng[3].bit_generator.state)
        arrivals = self.rng[2].poisson(cfg.arrival_rate, size=n_steps)
        cap = 2 * max_steps                                    # the reference pre-samples 2*max_steps sizes
        need = int(min(cap, int(arrivals.sum()))) if n_adm is None else int(min(cap, n_adm))
        # the reference always advances rng1/rng2 by 2*max_steps draws per reset; draw them all only when a
        # later stream-continuing reset could observe the difference (cheap prefix otherwise)
        cpu = np.around(self.rng[0].uniform(low=lo, high=hi, size=cap), decimals=2)[:need]
        mem = np.around(self.rng[1].uniform(low=lo, high=hi, size=cap), decimals=2)[:need]
        svc = self.rng[3].poisson(cfg.service_length, size=need) + 1
        return arrivals, cpu, mem, svc

    def rewind_to(self, cfg, steps_used: int, adm_used: int):
        s3, s4 = self._mark
        self.rng[2].bit_generator.state = s3
        self.rng[3].bit_generator.state = s4
        if steps_used:
            self.rng[2].poisson(cfg.arrival_rate, size=int(steps_used))
        if adm_used:
            self.rng[3].poisson(cfg.service_length, size=int(adm_used))


def pack_trace(per_env, n_steps: int):
    """per_env: list of (arrivals, cpu, mem, svc) -> (arrivals u16[N,T], admissions u32[N,J])."""
    N = len(per_env)
    J = max(1, max(len(t[1]) for t in per_env))
    arr = np.zeros((N, n_steps), np.uint16)
    adm = np.zeros((N, J), np.uint32)
    for i, (a, cpu, mem, svc) in enumerate(per_env):
        if a.max(initial=0) > 65535:
            raise ValueError("arrival count exceeds the u16 trace format")
        if svc.max(initial=0) > 65535:
            raise ValueError("service length exceeds the u16 trace/remaining format (service_length too large)")
        arr[i] = a
        cc = np.rint(cpu * 100).astype(np.uint32)
        mc = np.rint(mem * 100).astype(np.uint32)
        adm[i, : len(cc)] = cc | (mc << 8) | (svc.astype(np.uint32) << 16)
    return arr, adm


def sample_numpy_traces(cfg, streams, n_steps: int, n_adm: int | None, threads: int = 8):
    """`cfg`: one Config for all envs, or a list with one Config per env (sweeps over arrival_rate / service_length /
    sequence mapped onto the batch axis: in pre-sampled mode those parameters only shape the traces)."""
    cfgs = list(cfg) if isinstance(cfg, (list, tuple)) else [cfg] * len(streams)

    def one(i):
        return streams[i].draw(cfgs[i], n_steps, n_adm)
    if len(streams) > 4 and threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            per_env = list(ex.map(one, range(len(streams))))
    else:
        per_env = [one(i) for i in range(len(streams))]
    return pack_trace(per_env, n_steps)


def poisson_cdf_table(lam: float):
    """(kmin, thresholds u64[len]) with thresholds[i] = floor(P(X <= kmin+i) * 2^64), covering all mass that
    is representable at 2^-64; the last entry is forced to 2^64-1.  The kernel returns
    kmin + #{i : thresholds[i] <= u} for a uniform 64-bit u (exact inversion)."""
    lam = float(lam)
    if lam <= 0:
        return 0, np.array([np.iinfo(np.uint64).max], np.uint64)
    sd = math.sqrt(lam)
    kmin = max(0, int(math.floor(lam - 12 * sd - 12)))
    kmax = int(math.ceil(lam + 12 * sd + 40))
    ks = np.arange(0, kmax + 1, dtype=np.float64)
    logpmf = ks * math.log(lam) - lam - np.array([math.lgamma(k + 1.0) for k in ks])
    pmf = np.exp(logpmf)
    cdf = np.cumsum(pmf)
    cdf = np.minimum(cdf / cdf[-1], 1.0)
    sel = cdf[kmin:]
    # exact integer thresholds via python ints (float64 has 53 bits; fine for a sampling table)
    th = [min((1 << 64) - 1, int(c * 18446744073709551616.0)) if c < 1.0 else (1 << 64) - 1 for c in sel]
    th[-1] = (1 << 64) - 1
    # drop a redundant saturated tail
    while len(th) > 1 and th[-2] == (1 << 64) - 1:
        th.pop()
    return kmin, np.array(th, dtype=np.uint64)


def cdf_brackets(thresholds: np.ndarray, bits: int = 6) -> np.ndarray:
    """u16[2^bits + 1]: bracket[b] = #{i : thresholds[i] <= b << (64 - bits)}, bracket[2^bits] = len (search start)."""
    n = 1 << bits
    edges = (np.arange(n, dtype=np.uint64) << np.uint64(64 - bits))
    lo = np.searchsorted(thresholds, edges, side="right")
    return np.concatenate([lo, [len(thresholds)]]).astype(np.uint16)


def size_code_range(sequence: str):
    return SEQUENCE_CODES[sequence]
