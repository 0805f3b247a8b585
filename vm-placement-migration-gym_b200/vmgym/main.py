"""`python -m vmgym.main -a firstfit -e -c configs/10.yml` — the reference's entry point (main.py:31-87,89-112) on the
B200-native env: same arguments, same flow (seed everything, build env + agent, load or learn weights, evaluate, print
the Record summary).  `convex` needs cvxpy + SCIP (third-party MIP, not part of this package)."""
from __future__ import annotations

import argparse
import os
import random
from dataclasses import dataclass

import numpy as np
import torch
import yaml

from .config import Config
from .env import VmEnv


@dataclass
class Args:
    agent: str
    reward: str
    config: dict
    logdir: str | None = None
    output: str | None = None
    silent: bool = False
    jobname: str | None = None
    weightspath: str | None = None
    eval: bool = False
    debug: bool = False


def run(args: Args):
    config = args.config
    env_config = dict(config["environment"])
    env_config["reward_function"] = args.reward                      # main.py:34: the CLI reward overrides the YAML
    agent_config = dict(config.get("agents", {}).get(args.agent, {}))
    agent_config.pop("device", None)                                 # the device is the env's
    seed = env_config["seed"]
    torch.manual_seed(seed); random.seed(seed); np.random.seed(seed)  # main.py:40-45
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.set_float32_matmul_precision("high")
    # best-fit rows of the published tables need numpy's scalar-introsort tie order
    env = VmEnv(Config(**env_config), tiebreak=os.environ.get("VMGYM_TIEBREAK", "stable"))
    if args.agent == "firstfit":
        from .agents import FirstFitAgent
        agent = FirstFitAgent(env)
    elif args.agent == "bestfit":
        from .agents import BestFitAgent
        agent = BestFitAgent(env)
    elif args.agent == "ppo":
        from .ppo import PPOAgent, PPOConfig
        agent = PPOAgent(env, PPOConfig(**agent_config))
    elif args.agent == "drlvmp":
        from .drlvmp import DRLVMPAgent, DRLVMPConfig
        agent = DRLVMPAgent(env, DRLVMPConfig(**agent_config))
    else:
        raise SystemExit(f"Agent cannot be {args.agent}")
    if args.logdir and args.jobname:
        agent.set_log(jobname=args.jobname, logdir=args.logdir)
    if args.weightspath and os.path.exists(args.weightspath):
        agent.load_model(args.weightspath)
    else:
        agent.learn()
        if args.weightspath:
            os.makedirs(os.path.dirname(os.path.abspath(args.weightspath)), exist_ok=True)
            agent.save_model(args.weightspath)
    record = agent.test(show=not args.silent, output=args.output, debug=args.debug) if args.eval else None
    agent.end_log()
    return record


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-a", "--agent", required=True, choices=["ppo", "firstfit", "bestfit", "drlvmp"])
    ap.add_argument("-c", "--config", default="configs/10.yml")
    ap.add_argument("-r", "--reward", default="wr", choices=["wr", "ut", "kl"])
    ap.add_argument("-d", "--debug", action="store_true")
    ap.add_argument("-l", "--logdir")
    ap.add_argument("-j", "--jobname")
    ap.add_argument("-o", "--output", default="./output.json")
    ap.add_argument("-w", "--weightspath")
    ap.add_argument("-e", "--eval", action="store_true")
    ap.add_argument("-s", "--silent", default=False, action="store_true")
    a = ap.parse_args(argv)
    with open(a.config) as f:
        cfg = yaml.safe_load(f)
    return run(Args(agent=a.agent, reward=a.reward, config=cfg, logdir=a.logdir, output=a.output, silent=a.silent,
                    jobname=a.jobname, weightspath=a.weightspath, eval=a.eval, debug=a.debug))


if __name__ == "__main__":
    main()
