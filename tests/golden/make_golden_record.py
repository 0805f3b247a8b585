#!/usr/bin/env python
"""Golden vectors for the per-VM episode statistics of the reference's Record (src/record.py:34-96,110-134):
pending rates, slowdown rates and VM lifetimes of complete evaluation episodes, produced by the UNMODIFIED reference
classes (VmEnv in eval mode, FirstFit/BestFit agents, Base.record_testing_step, Record).  Build container only.

    python tests/golden/make_golden_record.py      ->  tests/golden/record.npz

Per case: cfg_json, agent, tiebreak, perturb, actions i16[T,V] (the exact action stream, so the oracle / CUDA replay
does not depend on the agents), pending f64[n], slowdown f64[m], lifetime i64[n], summary_json (Record.get_summary()).
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
    "rec_busy_firstfit": dict(base="10", over=dict(reward_function="wr", arrival_rate=0.35, service_length=40, eval_steps=1500,
                                                   seed=7), agent="firstfit"),
    "rec_busy_suspend": dict(base="10", over=dict(reward_function="ut", arrival_rate=0.5, service_length=25, eval_steps=1200,
                                                  seed=11, sequence="lowuniform"), agent="bestfit", perturb=0.2),
    "rec_p37_v70": dict(base="10", over=dict(pms=37, vms=70, reward_function="wr", arrival_rate=1.2, service_length=60,
                                             seed=5, eval_steps=900), agent="bestfit", perturb=0.05),
    "rec_s10_sparse": dict(base="10", over=dict(reward_function="wr", eval_steps=6000), agent="firstfit"),
    "rec_s100_bestfit": dict(base="100", over=dict(reward_function="wr", eval_steps=2300), agent="bestfit"),
    # `python main.py -a firstfit -e -c config/10.yml` (BASELINE configs[0]; CLI default reward wr) and the best-fit / ut
    # variant: full 100 000-step episodes with the real placement-matrix rank (SVD) — summaries only, no action stream
    "main_s10_firstfit_wr": dict(base="10", over=dict(reward_function="wr"), agent="firstfit", real_rank=True, keep_actions=False),
    "main_s10_bestfit_ut": dict(base="10", over=dict(reward_function="ut"), agent="bestfit", real_rank=True, keep_actions=False),
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
    if not spec.get("real_rank"):
        env._get_rank = lambda: 0                  # only feeds info['rank'] (env.py:317); skips the per-step SVD
    P, V, A = cfg["pms"], cfg["vms"], env.action_dim
    perturb = spec.get("perturb", 0.0)
    prng = np.random.default_rng(77)
    # Base.test (base.py:63-86) without tqdm / plots
    env.eval()
    agent.eval()
    obs, info = env.reset(seed=env.config.seed)
    done = False
    actions = []
    while not done:
        action = np.asarray(agent.act(obs)).astype(np.int64)
        if perturb:
            u = prng.random(V)
            action = np.where(u < perturb, P, action)              # suspend running VMs / keep waiting VMs waiting
        if spec.get("keep_actions", True):
            actions.append(action.astype(np.int16))
        obs, reward, done, truncated, info = env.step(action)
        agent.record_testing_step(reward, info)
    rec = agent.record
    summary = {k: (float(v) if np.ndim(v) == 0 else np.asarray(v, dtype=float).tolist()) for k, v in rec.get_summary().items()}
    out = {f"{name}.cfg_json": json.dumps(cfg), f"{name}.agent": spec["agent"], f"{name}.perturb": perturb,
           f"{name}.actions": np.array(actions, np.int16).reshape(-1, V),
           f"{name}.pending": np.array(rec.pending_rates, np.float64),
           f"{name}.slowdown": np.array(rec.slowdown_rates, np.float64),
           f"{name}.lifetime": np.array(rec.vm_lifetime, np.int64),
           f"{name}.summary_json": json.dumps(summary)}
    print(f"{name}: T={env.timestep - 1} vms={len(rec.pending_rates)} allocated={len(rec.slowdown_rates)} "
          f"avg pending {summary['average pending']} median {summary['median pending']} slowdown {summary['average slowdown']} "
          f"life {summary['average VM life']}", flush=True)
    return out


def main():
    out = {}
    for name, spec in CASES.items():
        out.update(run_case(name, spec))
    out["numpy_version"] = np.__version__
    path = os.path.join(HERE, "record.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
