from vmgym.ppo import PPOAgent, PPOConfig  # noqa: F401  (reference path: src/agents/ppo.py)
