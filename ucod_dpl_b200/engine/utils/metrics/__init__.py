from .metric import statistics  # noqa: F401
