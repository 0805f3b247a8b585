from vmgym.agents import FirstFitAgent  # noqa: F401  (reference path: src/agents/firstfit.py)
