from vmgym.agents import BestFitAgent  # noqa: F401  (reference path: src/agents/bestfit.py)
