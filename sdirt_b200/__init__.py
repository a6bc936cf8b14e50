from . import _engine  # noqa: F401
