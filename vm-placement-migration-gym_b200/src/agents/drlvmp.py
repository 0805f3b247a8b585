from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig  # noqa: F401  (reference path: src/agents/drlvmp.py)
