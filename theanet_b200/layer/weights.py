"""Parameter containers and initialisation (reference: theanet/layer/weights.py:10-81).

``Param`` plays the role of a Theano shared variable: it owns (a view of) device memory inside
the network's flat theta buffer once the net is assembled, and hands numpy copies to the host
through ``get_value`` -- the accessor the reference's ``borrow`` helper (weights.py:18-22) uses.
"""
import numpy as np


class Param:
    def __init__(self, value, name=''):
        self._host = np.ascontiguousarray(value, dtype=np.float32)
        self.shape = self._host.shape
        self.ndim = self._host.ndim
        self.size = self._host.size
        self.name = name
        self.tensor = None      # torch view into the flat theta buffer (set by NeuralNet)
        self.grad = None        # matching view into the flat gradient buffer
        self.vel = None         # matching view into the flat velocity buffer

    def bind(self, theta_view, vel_view, grad_view):
        import torch
        theta_view.copy_(torch.from_numpy(self._host).reshape(theta_view.shape))
        self.tensor, self.vel, self.grad = theta_view, vel_view, grad_view
        self._host = None

    def get_value(self, borrow=True):
        if self.tensor is None:
            return self._host
        return self.tensor.detach().cpu().numpy().reshape(self.shape)

    def set_value(self, value):
        value = np.ascontiguousarray(value, dtype=np.float32).reshape(self.shape)
        if self.tensor is None:
            self._host = value
        else:
            import torch
            self.tensor.copy_(torch.from_numpy(value).reshape(self.tensor.shape))

    def __str__(self):
        return self.name


def is_shared_var(x):
    return isinstance(x, Param)


def borrow(sharedvar, boro=True):
    return sharedvar.get_value(borrow=boro)


def init_wb(wb, rand_gen, size_w, size_b, fan_in, fan_out, actvn, name):
    """Same contract as the reference's init_wb (weights.py:25-81): ``wb`` is None (draw from
    ``rand_gen``), a pair of ndarrays (copy) or a pair of Params (share, used by TestVersion).

    4-D filters: random signs / sqrt(fan_in) (:51-54); matrices: U(-1,1)*sqrt(6/(fan_in+fan_out))
    (:56-57); sigmoid x4 (:62-63); bias +0.5 for softplus / relu / relu0N (:64-65).
    """
    if wb is not None and is_shared_var(wb[0]):
        return wb[0], wb[1]
    if wb is None:
        if len(size_w) == 4:
            w = (2. * rand_gen.randint(2, size=size_w) - 1) / np.sqrt(fan_in)
        else:
            w = rand_gen.uniform(low=-1, high=1, size=size_w) * np.sqrt(6 / (fan_in + fan_out))
        w = np.asarray(w, dtype=np.float32)
        b = np.zeros(size_b, dtype=np.float32)
        if actvn == 'sigmoid':
            w *= 4
        if actvn in ('softplus', 'relu') or actvn.startswith('relu0'):
            b += .5
    else:
        w, b = (np.asarray(t, dtype=np.float32) for t in wb)
    return Param(w, name + 'W'), Param(b, name + 'b')
