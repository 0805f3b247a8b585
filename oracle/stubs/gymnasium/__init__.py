"""Minimal stand-in for `gymnasium` (not installed in this image; no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference classes under
/root/reference (vmenv/envs/env.py:2-3,19,27-28; vmenv/__init__.py:1; src/agents/ppo.py:10,92)
can be imported by `tests/golden/make_golden.py` when the golden fixtures are (re)generated.
Only the handful of names those files touch are provided.
"""
import importlib

from . import spaces  # noqa: F401

_REGISTRY = {}


class Env:
    """The two members VmEnv relies on: a `reset` accepting seed/options, and `metadata`."""
    metadata = {}

    def reset(self, seed=None, options=None):
        return None


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = entry_point


def make(id, **kwargs):
    module_name, class_name = _REGISTRY[id].split(":")
    cls = getattr(importlib.import_module(module_name), class_name)
    return cls(**kwargs)
