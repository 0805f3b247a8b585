#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference classes.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py --only s10_firstfit

The reference imports `gymnasium` and `matplotlib`, which are not installed here; the two stand-ins under
oracle/stubs/ provide the handful of names it touches.  Nothing from the reference is copied:
this script drives `vmenv.envs.env.VmEnv`, `src.agents.firstfit.FirstFitAgent` and
`src.agents.bestfit.BestFitAgent` through their public API and records what they do.

Fixture layout (all arrays indexed by step t = 0..T-1 unless noted; "+1" arrays include the reset state at 0):
  cfg_json, seed, agent, tiebreak
  action i16[T,V], valid u8[T,V], reward f64[T], terminated u8[T]
  placement i16[T+1,V], vm_cpu_code u8[T+1,V], vm_mem_code u8[T+1,V]   (sizes are exact hundredths, env.py:212-219)
  cpu f64[T+1,P], memory f64[T+1,P]                                   (bit patterns of the fp64 accumulators)
  remaining_sum i64[T+1]  = sum_v remaining[v]*(v+1);  remaining_final i64[V];  suspended u8[T+1,V]
  counters i64[T,6] = timestep,total_requests,served,suspend_actions,place_actions,dropped  (after the step)
  scalars f64[T,5] = total_cpu_requested,total_memory_requested,waiting_ratio,target_cpu_mean,target_memory_mean
  obs_sha256 (hex of the float32 observation stream incl. the reset observation)
  mask_steps i32[K], mask_bits u8[K, ceil(V*A/8)]  (np.packbits of get_invalid_action_mask() BEFORE step mask_steps[k])
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VMGYM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), REF]
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402
import yaml  # noqa: E402

SCALAR_SORT_ENV = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2 FMA3"


class _StableArgsortNumpy:
    """Proxy for the `np` name inside src/agents/bestfit.py: argsort forced to kind='stable' (SURVEY §8c ruling i)."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def argsort(a, *args, **kw):
        kw["kind"] = "stable"
        return np.argsort(a, *args, **kw)


def base_cfg(name):
    return yaml.safe_load(open(os.path.join(REF, "config", f"{name}.yml")))["environment"]


CASES = {
    # name: (cfg overrides on config/<base>.yml, agent, tiebreak, steps, perturb prob, eval_mode, mask_every)
    "s10_firstfit_wr": dict(base="10", over=dict(reward_function="wr"), agent="firstfit", steps=2000, mask_every=100),
    "s10_firstfit_kl": dict(base="10", over=dict(reward_function="kl"), agent="firstfit", steps=2000),
    "s10_firstfit_ut": dict(base="10", over=dict(reward_function="ut"), agent="firstfit", steps=2000),
    "s10_bestfit_stable_ut": dict(base="10", over=dict(reward_function="ut"), agent="bestfit", tiebreak="stable",
                                  steps=3000),
    "s100_firstfit_wr": dict(base="100", over=dict(reward_function="wr"), agent="firstfit", steps=1000, mask_every=250),
    "s100_firstfit_kl": dict(base="100", over=dict(reward_function="kl"), agent="firstfit", steps=400),
    "s100_bestfit_stable_wr": dict(base="100", over=dict(reward_function="wr"), agent="bestfit", tiebreak="stable",
                                   steps=1500),
    "s100_bestfit_introsort_wr": dict(base="100", over=dict(reward_function="wr"), agent="bestfit",
                                      tiebreak="numpy_introsort", steps=1500),
    # busy small envs with adversarial action noise (out-of-range values, suspend storms, NULL-slot actions)
    "busy_adv_null": dict(base="10", over=dict(reward_function="kl", arrival_rate=0.35, service_length=40,
                                               training_steps=700, allow_null_action=True, seed=7),
                          agent="firstfit", steps=1500, perturb=0.15, mask_every=25, episodes=True),
    "busy_adv_nonull": dict(base="10", over=dict(reward_function="ut", arrival_rate=0.5, service_length=25, beta=0.3,
                                                 training_steps=10000, allow_null_action=False, seed=11,
                                                 sequence="lowuniform", cap_target_util=False),
                            agent="bestfit", tiebreak="stable", steps=1500, perturb=0.25, mask_every=25),
    "tiny_p3_v5_high": dict(base="10", over=dict(pms=3, vms=5, reward_function="kl", arrival_rate=0.3, service_length=12,
                                                 sequence="highuniform", seed=3, training_steps=10000),
                            agent="firstfit", steps=1200, perturb=0.2, mask_every=10),
    "odd_p37_v70_eval": dict(base="10", over=dict(pms=37, vms=70, reward_function="wr", arrival_rate=1.2,
                                                  service_length=60, seed=5, eval_steps=900, training_steps=300),
                             agent="bestfit", tiebreak="stable", steps=900, perturb=0.05, eval_mode=True, mask_every=90),
}


def run_case(name, spec):
    from vmenv.envs.env import VmEnv
    from vmenv.envs.config import Config
    from src.agents.firstfit import FirstFitAgent
    import src.agents.bestfit as bestfit_mod

    cfg = base_cfg(spec["base"])
    cfg.update(spec.get("over", {}))
    tiebreak = spec.get("tiebreak", "")
    if spec["agent"] == "bestfit":
        if tiebreak == "stable":
            bestfit_mod.np = _StableArgsortNumpy()
        else:
            bestfit_mod.np = np
            assert os.environ.get("NPY_DISABLE_CPU_FEATURES"), "introsort cases must run with SIMD sort dispatch disabled"
    env = VmEnv(Config(**cfg))
    agent = FirstFitAgent(env) if spec["agent"] == "firstfit" else bestfit_mod.BestFitAgent(env)
    if spec.get("eval_mode"):
        env.eval()
        env._get_rank = lambda: 0  # only feeds info['rank'] (env.py:317); skips the per-step SVD
    P, V, A = cfg["pms"], cfg["vms"], env.action_dim
    T = spec["steps"]
    perturb = spec.get("perturb", 0.0)
    mask_every = spec.get("mask_every", 0)
    prng = np.random.default_rng(20240917)

    obs, _ = env.reset(seed=cfg["seed"])
    sha = hashlib.sha256()
    sha.update(obs.tobytes())
    rec = {k: [] for k in ("action", "valid", "reward", "terminated", "placement", "vm_cpu_code", "vm_mem_code", "cpu",
                           "memory", "remaining_sum", "suspended", "counters", "scalars", "mask_steps", "mask_bits",
                           "reset_at", "reset_seed")}

    def snap():
        rec["placement"].append(env.vm_placement.astype(np.int16))
        for key, arr in (("vm_cpu_code", env.vm_cpu), ("vm_mem_code", env.vm_memory)):
            code = np.rint(arr * 100).astype(np.int64)
            assert np.array_equal(code / 100.0, arr), "sizes are not exact hundredths"
            rec[key].append(code.astype(np.uint8))
        rec["cpu"].append(env.cpu.copy())
        rec["memory"].append(env.memory.copy())
        rec["remaining_sum"].append(int(np.dot(env.vm_remaining_runtime.astype(np.int64), np.arange(1, V + 1))))
        rec["suspended"].append(env.vm_suspended.astype(np.uint8))

    snap()
    episode = 0
    for t in range(T):
        if mask_every and t % mask_every == 0:
            rec["mask_steps"].append(t)
            rec["mask_bits"].append(np.packbits(env.get_invalid_action_mask(True).reshape(-1)))
        action = np.asarray(agent.act(obs)).astype(np.int64)
        if perturb:
            u = prng.random(V)
            rnd = prng.integers(0, A + 2, size=V)                  # includes out-of-range A and A+1
            action = np.where(u < perturb / 2, rnd, action)
            action = np.where((u >= perturb / 2) & (u < perturb), P, action)   # suspend / stay-waiting storms
        obs, reward, term, trunc, info = env.step(action)
        sha.update(obs.tobytes())
        rec["action"].append(action.astype(np.int16))
        rec["valid"].append(np.asarray(info["valid"]).astype(np.uint8))
        rec["reward"].append(float(reward))
        rec["terminated"].append(int(term))
        rec["counters"].append([env.timestep, env.total_requests, env.served_requests, env.suspend_action,
                                env.place_action, env.dropped_requests])
        rec["scalars"].append([env.total_cpu_requested, env.total_memory_requested, env.waiting_ratio,
                               env.target_cpu_mean, env.target_memory_mean])
        snap()
        if term and spec.get("episodes") and t + 1 < T:
            # alternate the two reset flavours: reseeded (ppo.py:192) and stream-continuing (drlvmp.py:450-452 after seed())
            episode += 1
            if episode % 2 == 1:
                obs, _ = env.reset(seed=cfg["seed"] + episode)
                rec["reset_seed"].append(cfg["seed"] + episode)
            else:
                obs, _ = env.reset()
                rec["reset_seed"].append(-1)
            rec["reset_at"].append(t + 1)
            sha.update(obs.tobytes())
            # the post-reset state replaces the snapshot taken after the terminal step for index t+1
            for k in ("placement", "vm_cpu_code", "vm_mem_code", "cpu", "memory", "remaining_sum", "suspended"):
                rec[k].pop()
            snap()

    out = dict(
        cfg_json=json.dumps(cfg), seed=cfg["seed"], agent=spec["agent"], tiebreak=tiebreak,
        eval_mode=int(bool(spec.get("eval_mode"))),
        action=np.array(rec["action"], np.int16), valid=np.array(rec["valid"], np.uint8),
        reward=np.array(rec["reward"], np.float64), terminated=np.array(rec["terminated"], np.uint8),
        placement=np.array(rec["placement"], np.int16), vm_cpu_code=np.array(rec["vm_cpu_code"], np.uint8),
        vm_mem_code=np.array(rec["vm_mem_code"], np.uint8), cpu=np.array(rec["cpu"], np.float64),
        memory=np.array(rec["memory"], np.float64), remaining_sum=np.array(rec["remaining_sum"], np.int64),
        remaining_final=env.vm_remaining_runtime.astype(np.int64), suspended=np.array(rec["suspended"], np.uint8),
        counters=np.array(rec["counters"], np.int64), scalars=np.array(rec["scalars"], np.float64),
        obs_sha256=sha.hexdigest(), mask_steps=np.array(rec["mask_steps"], np.int32),
        mask_bits=np.array(rec["mask_bits"], np.uint8) if rec["mask_bits"] else np.zeros((0, 0), np.uint8),
        reset_at=np.array(rec["reset_at"], np.int64), reset_seed=np.array(rec["reset_seed"], np.int64),
        numpy_version=np.__version__,
    )
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: T={T} sha={sha.hexdigest()[:16]} sum_reward={np.sum(out['reward']):.9f} "
          f"counters={out['counters'][-1].tolist()} -> {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    for name, spec in CASES.items():
        if args.only and args.only != name:
            continue
        needs_scalar_sort = spec.get("tiebreak") == "numpy_introsort"
        if needs_scalar_sort and not os.environ.get("NPY_DISABLE_CPU_FEATURES"):
            # numpy picks its sort kernels at import time: re-exec with the SIMD sorts disabled
            env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=SCALAR_SORT_ENV)
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--only", name, "--child"], env=env)
            continue
        run_case(name, spec)


if __name__ == "__main__":
    main()
