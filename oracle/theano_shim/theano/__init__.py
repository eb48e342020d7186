"""Stand-in for the part of Theano that rakeshvar/theanet's training path uses.
TEST INFRASTRUCTURE ONLY -- see ../README.md.  Not Theano; not shipped; never on the product path."""
from ._core import config, shared, function, Var, SharedVar, Function  # noqa: F401
from . import tensor  # noqa: F401
from . import compile  # noqa: F401


def scan(*a, **k):
    raise NotImplementedError("theano_shim: scan (only used by the reference's HingeLayer) is not provided")
