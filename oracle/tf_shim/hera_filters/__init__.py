from . import dspec  # noqa: F401
