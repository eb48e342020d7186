from .._core import sigmoid, softplus, softmax, nnet_conv2d as conv2d  # noqa: F401
