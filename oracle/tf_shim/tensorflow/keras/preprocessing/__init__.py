from . import sequence  # noqa: F401
