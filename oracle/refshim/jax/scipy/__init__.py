from . import linalg, stats, special  # noqa: F401
