from . import norm  # noqa: F401
