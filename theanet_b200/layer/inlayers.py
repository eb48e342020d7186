"""InputLayer and ElasticLayer (reference: theanet/layer/inlayers.py:12-163)."""
import numpy as np

from .layer import Layer, Out


class InputLayer(Layer):
    def __init__(self, inpt, img_sz, num_maps=1, rand_gen=None):
        self.params = []
        self.inpt = inpt
        self.out_sz = img_sz
        self.num_maps = num_maps
        self.n_out = self.num_maps * self.out_sz ** 2
        self.representation = 'Input Maps:{} Sizes Input:{:2d} Output:{:2d}'.format(
            num_maps, img_sz, img_sz)
        self.output = Out(self, (num_maps, img_sz, img_sz))

    def TestVersion(self, inpt):
        return InputLayer(inpt, self.out_sz, self.num_maps)


class ElasticLayer(Layer):
    """Per-minibatch random distortion: translation, Gaussian-smoothed elastic field, zoom,
    rotation, nearest / bilinear resampling, pixel-flip noise (inlayers.py:29-155).  The random
    draws come from Philox streams keyed by (seed, step) instead of Theano's MT19937 streams;
    ``seed`` is drawn from ``rand_gen`` at the same point the reference seeds its RandomStreams
    (inlayers.py:72), so the numpy init stream stays aligned with the reference's."""

    def __init__(self, inpt, img_sz, num_maps=1, translation=0, zoom=1, magnitude=0, sigma=1,
                 pflip=0, angle=0, rand_gen=None, invert_image=False, nearest=False):
        self.inpt = inpt
        self.img_sz = img_sz
        self.translation = translation
        self.zoom = zoom
        self.magnitude = magnitude
        self.sigma = sigma
        self.pflip = pflip
        self.angle = angle
        self.invert = invert_image
        self.nearest = nearest
        self.out_sz = img_sz
        self.num_maps = num_maps
        self.n_out = self.num_maps * self.out_sz ** 2
        self.params = []
        self.representation = ('Elastic Maps:{:d} Size:{:2d} Translation:{:} Zoom:{} Mag:{:d} '
                               'Sig:{:d} Noise:{} Angle:{} Invert:{} Interpolation:{}'.format(
                                   self.num_maps, img_sz, translation, zoom, magnitude, sigma,
                                   pflip, angle, invert_image,
                                   'Nearest' if nearest else 'Linear'))
        assert zoom > 0
        self.output = Out(self, (num_maps, img_sz, img_sz))
        self.identity = (not (magnitude or translation or pflip or angle)) and zoom == 1
        self.has_grid = bool(magnitude or translation or angle or zoom != 1)
        self.seed = None
        self.filt = None
        if self.identity:
            return
        self.seed = int(rand_gen.randint(1e6)) if rand_gen is not None \
            else int(np.random.randint(1e6))
        if magnitude:
            # inlayers.py:87-91: float32 table truncated at +-sigma, divided by 2*pi*sigma^2
            var = sigma ** 2
            filt = np.array([[np.exp(-.5 * (i * i + j * j) / var)
                              for i in range(-sigma, sigma + 1)]
                             for j in range(-sigma, sigma + 1)], dtype=np.float32)
            filt /= np.float32(2 * np.pi * var)
            self.filt = filt
        # filled by the engine after each training step when debugging is on (inlayers.py:145-155)
        self.debugout = None

    def TestVersion(self, te_inpt):
        return ElasticLayer(te_inpt, self.img_sz, translation=0, zoom=1, magnitude=0, sigma=1,
                            pflip=0, angle=0, invert_image=self.invert, nearest=self.nearest)
