"""Dropout (reference: theanet/layer/dropout.py:9-31): Bernoulli(1-pdrop) mask, NO 1/(1-p)
rescale while training; the test twin multiplies by (1-pdrop)."""
import numpy as np

from .layer import Layer, Out


def draw_stream_seed(rand_gen):
    """The reference seeds a RandomStreams with rand_gen.randint(1e6) (dropout.py:10); we draw the
    same number at the same point and use it as the Philox key of this layer's mask stream."""
    return int(rand_gen.randint(1e6)) if rand_gen is not None else int(np.random.randint(1e6))


class DropOutLayer(Layer):
    def __init__(self, inpt, rand_gen=None, n_in=None, pdrop=0, test_scale=1.):
        self.seed = draw_stream_seed(rand_gen) if pdrop else None
        self.inpt = inpt
        self.params = []
        self.n_in, self.n_out = n_in, n_in
        self.pdrop = pdrop
        self.test_scale = test_scale     # (1 - pdrop) on the test twin, dropout.py:28-31
        self.output = Out(self, inpt.shape if inpt is not None else (n_in,))
        self.representation = "Drop:{:.0%} Out:{:3d}".format(pdrop, n_in)

    def TestVersion(self, inpt):
        return DropOutLayer(inpt, n_in=self.n_in, pdrop=0, test_scale=1 - self.pdrop)
