from .. import register  # noqa: F401
