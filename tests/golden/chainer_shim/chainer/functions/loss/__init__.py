from . import mean_squared_error  # noqa: F401
