from . import expr, plan  # noqa: F401
