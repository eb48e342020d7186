"""HiddenLayer (reference: theanet/layer/hidden.py:11-54): act(x.w + b), optional dropout."""
from .dropout import draw_stream_seed
from .layer import Layer, Out, activation_by_name
from .weights import init_wb


class HiddenLayer(Layer):
    def __init__(self, inpt, wts, rand_gen=None, n_in=None, n_out=None, pdrop=0, actvn='relu01',
                 reg=(), test_scale=1.):
        assert wts is not None or rand_gen is not None
        try:
            fan_in_out = n_in + n_out            # both fans, as the reference (hidden.py:21-27)
        except TypeError:
            fan_in_out = None
        self.w, self.b = init_wb(wts, rand_gen, (n_in, n_out), (n_out,), fan_in_out, fan_in_out,
                                 actvn, 'Hid')
        n_in, n_out = self.w.shape
        self.seed = draw_stream_seed(rand_gen) if pdrop else None     # dropout.py:10
        self.inpt = inpt
        self.params = [self.w, self.b]
        self.n_in, self.n_out = n_in, n_out
        self.actvn = actvn
        self.act = activation_by_name(actvn)
        self.pdrop = pdrop
        self.test_scale = test_scale
        self.reg = {"L1": 0, "L2": 0, "momentum": .95, "maxnorm": 0, "rate": 1}
        self.reg.update(reg)
        self.output = Out(self, (n_out,))
        self.representation = (
            "Hidden In:{:3d} Out:{:3d} Act:{} Drop%:{}\n\t  L1:{L1} L2:{L2} Momentum:{momentum} "
            "Max Norm:{maxnorm} Rate:{rate}".format(n_in, n_out, actvn, pdrop, **self.reg))

    def TestVersion(self, inpt):
        return HiddenLayer(inpt, (self.w, self.b), pdrop=0, actvn=self.actvn,
                           test_scale=1 - self.pdrop)
