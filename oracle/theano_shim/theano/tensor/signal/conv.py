from ..._core import signal_conv2d as conv2d  # noqa: F401
