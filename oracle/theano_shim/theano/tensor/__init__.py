from .._core import (  # noqa: F401
    tensor4, tensor3, matrix, vector, ivector, lvector, lscalar, iscalar, scalar,
    as_tensor_variable, constant, dot, exp, log, tanh, cos, sin, sqrt, sqr, abs_, maximum, minimum,
    clip, iround, cast, mean, sum, max, argmax, neq, eq, arange, stack, tensordot, concatenate,
    zeros_like, grad)
from . import nnet  # noqa: F401
from . import shared_randomstreams  # noqa: F401
from . import signal  # noqa: F401
