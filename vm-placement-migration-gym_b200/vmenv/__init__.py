"""Drop-in for the reference package `vmenv` (vmenv/__init__.py:1-6): registers "VmEnv-v1" when gymnasium is
installed and re-exports the B200-native env under the reference's import paths."""
try:  # gymnasium is optional here; the env does not depend on it
    from gymnasium.envs.registration import register

    register(id="VmEnv-v1", entry_point="vmenv.envs.env:VmEnv")
except Exception:  # pragma: no cover
    pass
