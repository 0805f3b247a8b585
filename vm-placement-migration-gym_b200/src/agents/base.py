from vmgym.agents import AgentBase as Base  # noqa: F401  (reference path: src/agents/base.py)
