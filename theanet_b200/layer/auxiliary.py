"""Auxiliary-input layers (reference: theanet/layer/auxiliary.py:14-160).

``LocationInfo`` turns a per-sample (2, 2) auxiliary input into n_aux[1] features through a random
mix of its two rows (training) or their mean (test) and two small dense layers (relu50, relu01).
``AuxConcatLayer`` appends those features to its input; it has no ``reg``, so -- exactly as in the
reference (layer.py:74-75) -- its LocationInfo weights are never updated.  ``SoftAuxLayer`` is a
softmax output layer whose scores get an extra linear term of the features; all of its eight
parameter tensors train with the layer's ``reg``.
"""
from .dropout import draw_stream_seed
from .hidden import HiddenLayer
from .layer import Layer, Out, activation_by_name
from .outlayers import OutputLayer
from .weights import init_wb


class LocationInfo():
    def __init__(self, wts, rand_gen=None, n_aux=(5, 9), boost=1, test_version=False):
        # the reference creates the random stream first (auxiliary.py:25), then the weights
        self.seed = None if test_version else draw_stream_seed(rand_gen)
        self.test_version = test_version
        self.boost = boost
        n_aux_hid, n_aux_out = n_aux
        self.n_hid, self.n_out = n_aux_hid, n_aux_out
        self.act1, self.act2 = activation_by_name("relu50"), activation_by_name("relu01")
        loc1_wts = None if wts is None else wts[:2]
        self.w1, self.b1 = init_wb(loc1_wts, rand_gen, (2, n_aux_hid), n_aux_hid, n_aux_hid + 2,
                                   n_aux_hid + 2, "relu50", 'Loc1')
        loc2_wts = None if wts is None else wts[2:]
        self.w2, self.b2 = init_wb(loc2_wts, rand_gen, (n_aux_hid, n_aux_out), n_aux_out,
                                   n_aux_out + n_aux_hid, n_aux_out + n_aux_hid, "relu01", 'Loc2')
        self.aux_inpt = Out(self, (2, 2))
        self.output = Out(self, (n_aux_out,))
        self.params = [self.w1, self.b1, self.w2, self.b2]


AUX_TYPES = {'LocationInfo': LocationInfo}


class AuxConcatLayer(Layer):
    def __init__(self, inpt, wts, rand_gen, n_in, n_aux, aux_type, boost=1, test_version=False):
        self.aux_info = AUX_TYPES[aux_type](wts, rand_gen, n_aux=n_aux, boost=boost,
                                            test_version=test_version)
        self.aux_inpt = self.aux_info.aux_inpt
        self.inpt = inpt
        self.n_aux = n_aux
        self.n_in = n_in
        self.n_out = n_aux[-1] + n_in
        self.aux_type = aux_type
        self.boost = boost
        self.params = self.aux_info.params
        self.output = Out(self, (self.n_out,))
        self.representation = "AuxConcat In:{:3d} Aux:{} Out:{:3d} ".format(n_in, n_aux, self.n_out)

    def TestVersion(self, te_inpt):
        return AuxConcatLayer(te_inpt, self.params, None, self.n_in, self.n_aux, self.aux_type,
                              boost=self.boost, test_version=True)


class SoftAuxLayer(HiddenLayer, OutputLayer):
    def __init__(self, inpt, wts, rand_gen, n_in, n_out, n_aux, aux_type, reg=(), loss="nll",
                 boost=1, test_version=False):
        hidden_wts = None if wts is None else wts[:2]
        HiddenLayer.__init__(self, inpt, hidden_wts, rand_gen, n_in, n_out, actvn='linear',
                             reg=reg, pdrop=0)
        aux_wts = None if wts is None else wts[2:6]
        self.aux_info = AUX_TYPES[aux_type](aux_wts, rand_gen, n_aux=n_aux, boost=boost,
                                            test_version=test_version)
        cross_wts = None if wts is None else wts[6:]
        n_aux_hid, n_aux_out = n_aux
        self.cross_w, self.cross_b = init_wb(cross_wts, rand_gen, (n_aux_out, n_out), n_out,
                                             n_aux_out + n_out, n_aux_out + n_out, 'softmax',
                                             'SoftAuxCross')
        self.aux_inpt = self.aux_info.aux_inpt
        self.n_aux = n_aux
        self.n_out = n_out
        self.aux_type = aux_type
        self.boost = boost
        self.loss = loss
        self.params = self.params + self.aux_info.params + [self.cross_w, self.cross_b]
        self.representation = (
            "SoftAux In:{:3d} Aux:{} Out:{:3d}\n\t  L1:{L1} L2:{L2} Momentum:{momentum} "
            "Max Norm:{maxnorm} Rate:{rate}".format(n_in, n_aux, n_out, **self.reg))
        self._handles()
        self.features = self.logprob
        self.probs = self.output
        self.kind = 'SOFTMAX'
        if not test_version:
            self.cost()

    def TestVersion(self, inpt):
        return SoftAuxLayer(inpt, self.params, rand_gen=None, n_in=self.n_in, n_out=self.n_out,
                            n_aux=self.n_aux, aux_type=self.aux_type, boost=self.boost,
                            test_version=True)
