from .._core import RandomStreams  # noqa: F401
