from ..._core import pool_2d  # noqa: F401
