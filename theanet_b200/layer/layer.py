"""Layer base class, activation registry and tensor handles (reference: theanet/layer/layer.py).

A layer here is host-side metadata (shapes, hyper-parameters, parameter handles) plus the
information the engine in ``theanet_b200.neuralnet`` needs to launch its forward / backward
kernels; ``.output`` is a :class:`Out` handle that resolves to a device tensor once the network
has allocated its buffers, not a symbolic expression.
"""
from .. import _C
from .weights import borrow

ACTIVATION_NAMES = (['sigmoid', 'softplus', 'softmax', 'linear', 'scaled_tanh', 'relu', 'tanh'] +
                    ['relu{:02d}'.format(i) for i in range(100)])


class Activation:
    """Named activation (layer.py:11-24); ``code``/``nn`` are what the kernels take."""

    def __init__(self, name):
        self.name = name
        if name in ('softmax', 'Softmax'):
            self.code, self.nn = _C.ACT_LINEAR, 0        # the softmax itself is a separate kernel
        else:
            self.code, self.nn = _C.act_code(name)

    def __str__(self):
        return self.name


def activation_by_name(name):
    if name in ('Softmax', 'softmax') or name in ACTIVATION_NAMES:
        return Activation(name)
    raise NotImplementedError("Unknown Activation Specified: " + name)


class Out:
    """Handle on a layer's output: who produces it and its per-sample shape."""

    def __init__(self, layer, shape):
        self.layer = layer
        self.shape = tuple(shape)       # per-sample shape, e.g. (C, S, S) or (n,)
        self.tensor = None              # device tensor (B, *shape), set by the engine

    def flatten(self, ndim=2):
        assert ndim == 2
        n = 1
        for s in self.shape:
            n *= s
        f = Out(self.layer, (n,))
        f.parent = self
        return f


class Layer:
    params = []

    def __str__(self):
        return self.representation

    def get_wts(self):
        return [borrow(p) for p in self.params]

    def get_updates(self, cost=None, rate=None):
        """The reference builds per-parameter Theano update pairs here (layer.py:70-107); in this
        implementation the whole rule runs in one fused kernel over the flat buffers
        (tn_sgd_momentum_maxnorm_update), so this only reports which tensors take part."""
        if not hasattr(self, "reg") or not self.reg['rate']:
            return []
        return list(self.params)

    def update_segments(self):
        """(param, reg) pairs for the fused optimiser."""
        if not hasattr(self, "reg"):
            return []
        return [(p, self.reg) for p in self.params]
