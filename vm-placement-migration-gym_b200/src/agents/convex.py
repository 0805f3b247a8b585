from vmgym.convex import ConvexAgent, ConvexConfig  # noqa: F401  (reference path: src/agents/convex.py)
