"""ColorLayer (reference: theanet/layer/color.py:9-52): random white balance and gamma curves per
(sample, map) while training; the test twin is the identity.  The three U(-1,1) draws per
(sample, map) come from the layer's Philox stream (tn_color_jitter) instead of Theano's
RandomStreams; the seed is drawn from ``rand_gen`` where the reference seeds its stream (:30)."""
import numpy as np

from .dropout import draw_stream_seed
from .layer import Layer, Out


class ColorLayer(Layer):
    def __init__(self, inpt, img_sz, num_maps=3, rand_gen=None, balance=1, gamma=1, maxval=1):
        self.params = []
        self.inpt = inpt
        self.out_sz = img_sz
        self.num_maps = num_maps
        self.n_out = self.num_maps * self.out_sz ** 2
        self.balance, self.gamma, self.maxval = balance, gamma, maxval
        self.representation = 'Color Maps:{} Size:{:2d} Balance:{:.2f} Gamma:{:.2f} Maxval:{}'.format(
            num_maps, img_sz, balance, gamma, maxval)
        self.output = Out(self, (num_maps, img_sz, img_sz))
        self.identity = gamma == 1 and balance == 1          # :26-28 (maxval is ignored then)
        self.seed = None
        if self.identity:
            return
        assert gamma > 0 and balance > 0
        self.seed = draw_stream_seed(rand_gen)
        # np.log(a) reaches the graph as a floatX constant (color.py:33)
        self.log_balance = float(np.float32(np.log(balance)))
        self.log_gamma = float(np.float32(np.log(gamma)))

    def TestVersion(self, inpt):
        return ColorLayer(inpt, self.out_sz, num_maps=self.num_maps, rand_gen=None, balance=1,
                          gamma=1, maxval=1)
