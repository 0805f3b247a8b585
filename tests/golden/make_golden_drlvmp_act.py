#!/usr/bin/env python
"""Golden vectors for the WHOLE DRLVMPAgent.act loop (src/agents/drlvmp.py:504-530: for every waiting VM in slot order the
network picks one of four heuristics on the working observation, whose placement entry is then overwritten), recorded from
the UNMODIFIED reference agent with a randomly initialised network.  Build container only (needs /root/reference).

    python tests/golden/make_golden_drlvmp_act.py   ->  tests/golden/drlvmp_act.npz

Per case `<c>.`: cfg_json, hidden, `sd.<key>` (the network's state dict incl. the NoisyNet noise buffers), obs f32[n, D]
(observations with waiting VMs, taken from a reference run), action i64[n, V] (agent.act(obs)), choices i8[n, V] (the heuristic
index the network picked for each waiting VM, -1 elsewhere) and margin f32[n, V] (best q minus second-best q at that decision).
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VMGYM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), REF]
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")   # torch.compile(self.dqn) then runs the module eagerly (the CPU inductor toolchain is incomplete here)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

CASES = {
    "act_s10": dict(base="10", over=dict(reward_function="wr", arrival_rate=0.4, service_length=40), hidden=64, steps=260, sample=(40, 90, 140, 200, 259)),
    "act_s100": dict(base="100", over=dict(reward_function="wr", service_length=100), hidden=64, steps=420, sample=tuple(range(120, 420, 12))),
}


def run_case(name, spec):
    from vmenv.envs.env import VmEnv
    from vmenv.envs.config import Config
    from src.agents.firstfit import FirstFitAgent
    import src.agents.drlvmp as ref
    cfg = yaml.safe_load(open(os.path.join(REF, "config", f"{spec['base']}.yml")))["environment"]
    cfg.update(spec["over"])
    torch.manual_seed(7); random.seed(7); np.random.seed(7)
    env = VmEnv(Config(**cfg))
    agent = ref.DRLVMPAgent(env, ref.DRLVMPConfig(hidden_size=spec["hidden"], memory_size=256, batch_size=8))
    agent.eval()
    ff = FirstFitAgent(env)
    obs, _ = env.reset(seed=cfg["seed"])
    rng = np.random.default_rng(5)
    V = cfg["vms"]
    rows_obs, rows_act, rows_choice, rows_margin = [], [], [], []
    for t in range(spec["steps"]):
        a = ff.act(obs)
        a = np.where(rng.random(a.size) < 0.5, env.vm_placement, a)       # leave half of the proposals unplaced: waiting VMs that fit
        obs, *_ = env.step(a)
        if t in spec["sample"]:
            # observe the network's decisions from outside: wrap _select_action on the INSTANCE
            log = []
            orig = agent._select_action

            def sel(o, orig=orig, log=log):
                with torch.no_grad():
                    q = agent.dqn(o).flatten()
                top = torch.topk(q, 2).values
                log.append((int(q.argmax()), float(top[0] - top[1])))
                return orig(o)
            agent._select_action = sel
            with torch.no_grad():
                action = agent.act(obs.copy())
            agent._select_action = orig
            waiting = np.flatnonzero(obs[:V] == env.WAIT_STATUS)
            assert len(log) == len(waiting)
            ch = np.full(V, -1, np.int8); mg = np.zeros(V, np.float32)
            for v, (c, m) in zip(waiting, log):
                ch[v], mg[v] = c, m
            # torch.argsort / argmin leave the order of equal keys unspecified (like np.argsort in bestfit.py): keep only the
            # observations in which no decision hinges on a tie (the PM loads in the working observation never change during act,
            # drlvmp.py:557-565, so keys and fits are those of the original observation)
            P = cfg["pms"]
            cpu, mem = obs[3 * V:3 * V + P], obs[3 * V + P:]
            tie = False
            for v in waiting:
                vc, vm = obs[V + v], obs[2 * V + v]
                if ch[v] in (0, 3):
                    fit = (cpu + vc <= 1) & (mem + vm <= 1)
                    key = (cpu + mem).astype(np.float32)
                    pm = int(action[v])
                    if pm < P and np.count_nonzero(fit & (key == key[pm])) > 1:
                        tie = True
                else:
                    key = (cpu * vc + mem * vm) if ch[v] == 1 else np.sqrt((cpu - vc) ** 2 + (mem - vm) ** 2)
                    if np.count_nonzero(key == key.min()) > 1:
                        tie = True
            if tie or len(waiting) == 0:
                continue
            rows_obs.append(obs.copy()); rows_act.append(np.asarray(action, np.int64)); rows_choice.append(ch); rows_margin.append(mg)
    sd = {k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k: v.detach().numpy() for k, v in agent.dqn.state_dict().items()}
    out = {f"{name}.cfg_json": json.dumps(cfg), f"{name}.hidden": spec["hidden"], f"{name}.obs": np.array(rows_obs, np.float32),
           f"{name}.action": np.array(rows_act), f"{name}.choices": np.array(rows_choice), f"{name}.margin": np.array(rows_margin)}
    for k, v in sd.items():
        out[f"{name}.sd.{k}"] = v
    n_dec = int((np.array(rows_choice) >= 0).sum())
    mg = np.array(rows_margin)[np.array(rows_choice) >= 0]
    print(f"{name}: {len(rows_obs)} observations, {n_dec} decisions, choice histogram {np.bincount(np.array(rows_choice)[np.array(rows_choice) >= 0], minlength=4).tolist()}, "
          f"min margin {mg.min():.3e}, 5th percentile {np.percentile(mg, 5):.3e}", flush=True)
    return out


def main():
    out = {}
    for name, spec in CASES.items():
        out.update(run_case(name, spec))
    path = os.path.join(HERE, "drlvmp_act.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
