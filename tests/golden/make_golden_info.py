#!/usr/bin/env python
"""Golden vectors for the eval-mode `info` dict and the two host-side logs of the reference env
(vmenv/envs/env.py:203,205,289,293,298-318): `vm_planned_runtime`, `vm_arrival_steps`, and what
Base.record_testing_step (src/agents/base.py:131-149) reads every step.  Produced by the UNMODIFIED reference classes;
build container only.

    python tests/golden/make_golden_info.py      ->  tests/golden/info.npz

Per case: cfg_json, actions i16[T,V], info_keys (json list, order of env.py:299-317 + step's action/valid),
timestep i64[T+1] (info["timestep"] of reset and of every step), planned i64[T+1,V] (vm_planned_runtime),
arrival_flat i64[n] + arrival_off i64[V+1] (the final vm_arrival_steps lists, ragged), used_pm i64[T]
(record.used_pm, base.py:134), served i64[T], rank i64[T] (real matrix_rank), waiting_ratio f64[T].
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VMGYM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), REF, HERE]
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402
from make_golden import _StableArgsortNumpy, base_cfg  # noqa: E402

CASES = {
    "info_busy_suspend": dict(base="10", over=dict(reward_function="ut", arrival_rate=0.5, service_length=25, eval_steps=700,
                                                   seed=11, sequence="lowuniform"), agent="bestfit", perturb=0.2),
    "info_p37_v70": dict(base="10", over=dict(pms=37, vms=70, reward_function="wr", arrival_rate=1.2, service_length=60,
                                              seed=5, eval_steps=400), agent="firstfit", perturb=0.05),
}


def run_case(name, spec):
    from vmenv.envs.env import VmEnv
    from vmenv.envs.config import Config
    from src.agents.firstfit import FirstFitAgent
    import src.agents.bestfit as bestfit_mod
    cfg = base_cfg(spec["base"])
    cfg.update(spec.get("over", {}))
    bestfit_mod.np = _StableArgsortNumpy()
    env = VmEnv(Config(**cfg))
    agent = FirstFitAgent(env) if spec["agent"] == "firstfit" else bestfit_mod.BestFitAgent(env)
    P, V = cfg["pms"], cfg["vms"]
    perturb = spec.get("perturb", 0.0)
    prng = np.random.default_rng(99)
    env.eval()
    agent.eval()
    obs, info = env.reset(seed=env.config.seed)
    timestep, planned = [info["timestep"]], [env.vm_planned_runtime.copy()]
    actions, used_pm, served, rank, wr = [], [], [], [], []
    keys = None
    done = False
    while not done:
        action = np.asarray(agent.act(obs)).astype(np.int64)
        u = prng.random(V)
        action = np.where(u < perturb, P, action)
        actions.append(action.astype(np.int16))
        obs, reward, done, truncated, info = env.step(action)
        agent.record_testing_step(reward, info)                     # the reference consumer of info (base.py:131-149)
        keys = list(info.keys())
        timestep.append(info["timestep"])
        planned.append(env.vm_planned_runtime.copy())
        used_pm.append(agent.record.used_pm[-1])
        served.append(agent.record.served_requests[-1])
        rank.append(int(info["rank"]))
        wr.append(float(info["waiting_ratio"]))
    arr = env.vm_arrival_steps
    off = np.zeros(V + 1, np.int64)
    off[1:] = np.cumsum([len(a) for a in arr])
    flat = np.array([x for a in arr for x in a], np.int64)
    print(f"{name}: T={len(actions)} arrivals={flat.size} keys={len(keys)}", flush=True)
    return {f"{name}.cfg_json": json.dumps(cfg), f"{name}.actions": np.array(actions, np.int16),
            f"{name}.info_keys": json.dumps(keys), f"{name}.timestep": np.array(timestep, np.int64),
            f"{name}.planned": np.array(planned, np.int64), f"{name}.arrival_flat": flat, f"{name}.arrival_off": off,
            f"{name}.used_pm": np.array(used_pm, np.int64), f"{name}.served": np.array(served, np.int64),
            f"{name}.rank": np.array(rank, np.int64), f"{name}.waiting_ratio": np.array(wr, np.float64)}


def main():
    out = {}
    for name, spec in CASES.items():
        out.update(run_case(name, spec))
    out["numpy_version"] = np.__version__
    path = os.path.join(HERE, "info.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
