"""The reference's experiment drivers mapped onto the batch axis.

The reference runs one OS process per (agent, load, service length, VM-size mix, seed) point, 8 at a time
(exp.py:1, exp_suspension.py:75-108, exp_vm_size.py:13-60, exp_performance.py:26-83), each a 100 000-step Python
episode.  Here every point is one env of a single VecVmEnv batch: the traces are pre-sampled per env from that env's
own parameters exactly as the reference samples them (numpy PCG64, env.py:172-226), the fused agent + step kernel
runs all points together, and the episode statistics (Record) are accumulated on the device.  Rows come out in the
reference's CSV column formats, digit for digit for the heuristic agents (tests/test_env_cuda.py).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .config import Config
from .vec_env import VecVmEnv

# best-fit rows of the published tables were produced on numpy's scalar introsort path (SURVEY §8c ruling ii)
PUBLISHED_TIEBREAK = {"firstfit": "stable", "bestfit": "numpy_introsort"}


def run_sweep(base: dict | Config, points: list[dict], agent: str, tiebreak: str | None = None, device="cuda",
              vm_stats: bool = True, chunk: int = 1000):
    """Evaluate `agent` ("firstfit" | "bestfit") on one env per point; a point overrides arrival_rate / service_length /
    sequence / seed of `base`.  Returns one dict per point: the point's overrides plus VecVmEnv.evaluate's columns."""
    base_d = dataclasses.asdict(base) if isinstance(base, Config) else dict(base)
    cfgs = [Config(**{**base_d, **pt}) for pt in points]
    vec = VecVmEnv(Config(**base_d), len(cfgs), device=device, rng="numpy", env_configs=cfgs,
                   tiebreak=tiebreak or PUBLISHED_TIEBREAK[agent])
    if vm_stats:
        vec.enable_vm_stats()
    res = vec.evaluate(agent, chunk=chunk)
    vec.close()
    return [{**pt, **{k: v[i] for k, v in res.items()}} for i, pt in enumerate(points)]


def suspension_points(pms: int = 100, service_lengths=None, loads=None):
    """The grid of exp_suspension.py:75-85 (committed script: service lengths 100..3900 step 200 at load 1, loads
    0.2..1.0 at service length 1000; the published data/exp_suspension/data.csv used 100..4100 step 1000 and loads
    0.5..1.1); arrival_rate = round(pms / 0.55 / service_length * load, 3) (exp_suspension.py:19)."""
    service_lengths = np.arange(100, 4100, 200) if service_lengths is None else service_lengths
    loads = np.arange(0.2, 1.1, 0.1) if loads is None else loads
    pts = [(1.0, int(sr)) for sr in service_lengths] + [(float(load), 1000) for load in loads]
    return [dict(service_length=sr, arrival_rate=float(np.round(pms / 0.55 / sr * load, 3)), _load=load) for load, sr in pts]


def suspension_rows(base: dict, agent: str, points=None, **kw):
    """exp_suspension.py:50-60 rows: 'Agent, Load, Service Length, Total Served, Valid Suspend Actions, Valid Actions,
    Life, Average Pending, Average Slowdown, Max Slowdown'."""
    base = {**base, "reward_function": "wr", "sequence": "uniform"}                  # exp_suspension.py:15-17
    points = suspension_points(base["pms"]) if points is None else points
    loads = [pt["_load"] for pt in points]
    res = run_sweep(base, [{k: v for k, v in pt.items() if not k.startswith("_")} for pt in points], agent, **kw)
    rows = []
    for load, r in zip(loads, res):
        rows.append("%s,%.1f,%d,%d,%d,%d,%d,%.3f,%.3f,%.3f" % (
            agent, load, r["service_length"], r["total served VMs"], r["total suspend actions"],
            r["total suspend actions"] + r["total place actions"], r["average VM life"], r["average pending"],
            r["average slowdown"], r["max slowdown"]))
    return rows


def seed_mean_row(base: dict, agent: str, seeds, label: str | None = None, **kw):
    """One row of exp_vm_size.py:62-96 / exp_performance.py-style tables: 5 seeds of one configuration averaged:
    'Model, Return, Drop Rate, Served VM, Suspend Actions, CPU Mean, CPU Variance, Memory Mean, Memory Variance,
    Waiting Ratio' (returns are rounded to 3 decimals per seed first, record.py:108)."""
    res = run_sweep(base, [dict(seed=int(s)) for s in seeds], agent, vm_stats=False, **kw)
    m = lambda k: float(np.mean([r[k] for r in res]))                                  # noqa: E731
    ret = float(np.mean([np.round(r["total rewards"], 3) for r in res]))
    return "%s,%.4f,%.4f,%d,%d,%.4f,%.4f,%.4f,%.4f,%.4f" % (
        label or agent, ret, m("drop rate"), m("total served VMs"), m("total suspend actions"), m("cpu mean"), m("cpu var"),
        m("memory mean"), m("memory var"), m("waiting ratio"))


def vm_size_rows(base: dict, agent: str, seeds=range(5), **kw):
    """exp_vm_size.py:13-20,98-108: lowuniform at arrival pms/0.375/service_length, highuniform at pms/0.625/service_length."""
    rows = []
    for seq, frac in (("lowuniform", 0.375), ("highuniform", 0.625)):
        b = {**base, "sequence": seq, "arrival_rate": base["pms"] / frac / base["service_length"]}
        rows.append(seed_mean_row(b, agent, seeds, **kw))
    return rows


def performance_row(base: dict, agent: str, load: float, reward: str, seeds=range(5), label: str | None = None, **kw):
    """One row of exp_performance.py:20-147 / exp_performance_small.py: `seeds` runs of one (agent, load) point averaged:
    'Agent, Load, Return, Drop Rate, Served VM, Suspend Actions, CPU Mean, CPU Variance, Memory Mean, Memory Variance, Pending
    Rate, Waiting Ratio, Slowdown Rate'.  arrival_rate = round(pms / 0.55 / service_length * load, 4) (:26).
    The reference's Memory Variance column is the variance ACROSS THE RUNS of each (step, PM) entry (`np.var(memory, axis=0)`,
    :117 — the CPU column uses axis=2, across PMs); it couples the runs step by step and is not accumulated here: the column
    printed is the across-PM variance like the CPU one."""
    b = {**base, "reward_function": reward,
         "arrival_rate": float(np.round(base["pms"] / 0.55 / base["service_length"] * load, 4))}
    res = run_sweep(b, [dict(seed=int(s)) for s in seeds], agent, **kw)
    m = lambda k: float(np.mean([r[k] for r in res]))                                  # noqa: E731
    ret = float(np.mean([np.round(r["total rewards"], 3) for r in res]))               # record.py:108 rounds per run
    return "%s,%.2f,%.3f,%.3f,%d,%d,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f" % (
        label or agent, load, ret, m("drop rate"), m("total served VMs"), m("total suspend actions"), m("cpu mean"), m("cpu var"),
        m("memory mean"), m("memory var"), m("average pending"), m("waiting ratio"), m("average slowdown"))
