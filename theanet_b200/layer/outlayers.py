"""Output layers and their losses (reference: theanet/layer/outlayers.py:12-147).

SoftmaxLayer with 'nll' is the hot path (fused head, head.cu); 'nllsq', the truncated 'nllNN'
losses, ExpLossLayer and HingeLayer run through tn_output_loss_fwd_bwd / tn_output_test_stats
(softmax.cu).  CenteredOutLayer and the scan-based hinge_max are not implemented."""
import numpy as np

from .. import _C
from .hidden import HiddenLayer
from .layer import Out

OUT_KINDS = {'SOFTMAX': _C.OUT_SOFTMAX, 'ExpLoss': _C.OUT_EXPLOSS, 'Hinge': _C.OUT_HINGE}


class OutputLayer(object):
    def cost(self, y=None):
        """The loss the engine compiles in, as (TN_LOSS_* code, log threshold) -- the dispatch of
        outlayers.py:12-36, including its fallback to plain NLL for an unreadable 'nllXX'."""
        if self.loss == "nll":
            return _C.LOSS_NLL, 0.0
        elif self.loss == "nllsq":
            return _C.LOSS_NLLSQ, 0.0
        elif self.loss.startswith("nll"):
            try:
                threshold = int(self.loss[-2:]) / 100
                threshold = np.clip(threshold, 0, 1)
            except ValueError:
                print("Did not understand {}, using plain NLL".format(self.loss))
                threshold = 1.0
            print("Using threshold: ", threshold)
            with np.errstate(divide='ignore'):
                return _C.LOSS_NLLTRUNC, float(np.log(threshold))       # a negative number
        elif self.loss == "hinge":
            return _C.LOSS_HINGE, 0.0
        elif self.loss == "exp":
            return _C.LOSS_EXP, 0.0
        raise NotImplementedError("Loss : " + self.loss)

    def features_and_predictions(self):
        return self.features, self.y_preds

    def sym_and_oth_err_rate(self, y=None):
        return "sym_err_rate", "second_stat"

    def _handles(self):
        # resolved by the engine: features, log-probabilities, argmax
        self.logprob = Out(self, (self.n_out,))
        self.y_preds = Out(self, ())


class SoftmaxLayer(HiddenLayer, OutputLayer):
    def __init__(self, inpt, wts, rand_gen=None, n_in=None, n_out=None, reg=(), loss="nll"):
        HiddenLayer.__init__(self, inpt, wts, rand_gen, n_in, n_out, actvn='Softmax', reg=reg,
                             pdrop=0)
        self._handles()
        self.features = self.logprob            # outlayers.py:93
        self.probs = self.output
        self.kind = 'SOFTMAX'
        self.loss = loss
        if loss is not None:
            self.cost()
        self.representation = (
            "Softmax In:{:3d} Out:{:3d} Loss:{}\n\t  L1:{L1} L2:{L2} Momentum:{momentum} "
            "Max Norm:{maxnorm} Rate:{rate}".format(self.n_in, self.n_out, self.loss, **self.reg))

    def TestVersion(self, inpt):
        return SoftmaxLayer(inpt, (self.w, self.b), loss=None)


class ExpLossLayer(HiddenLayer, OutputLayer):
    """Scores centred per row, loss mean exp(-o[y]) (outlayers.py:105-126)."""

    def __init__(self, inpt, wts, rand_gen=None, n_in=None, n_out=None, reg=()):
        HiddenLayer.__init__(self, inpt, wts, rand_gen, n_in, n_out, actvn='linear', reg=reg,
                             pdrop=0)
        self._handles()
        self.features = Out(self, (self.n_out,))        # the centred scores (:117)
        self.probs = self.output
        self.kind = 'ExpLoss'
        self.loss = 'exp'
        self.representation = (
            "ExpLoss In:{:3d} Out:{:3d} Loss:{}\n\t  L1:{L1} L2:{L2} Momentum:{momentum} "
            "Max Norm:{maxnorm} Rate:{rate}".format(self.n_in, self.n_out, self.loss, **self.reg))

    def TestVersion(self, inpt):
        return ExpLossLayer(inpt, (self.w, self.b))


class HingeLayer(HiddenLayer, OutputLayer):
    """Raw scores are features, 'logprob' and 'probs' at once; multi-class hinge loss averaged
    over every (sample, class) pair (outlayers.py:62-64,129-147)."""

    def __init__(self, inpt, wts, rand_gen=None, n_in=None, n_out=None, reg=()):
        HiddenLayer.__init__(self, inpt, wts, rand_gen, n_in, n_out, actvn='linear', reg=reg,
                             pdrop=0)
        self._handles()
        self.features = self.logprob
        self.probs = self.output
        self.kind = 'Hinge'
        self.loss = 'hinge'
        self.representation = (
            "SVM In:{:3d} Out:{:3d} Loss:{}\n\t  L1:{L1} L2:{L2} Momentum:{momentum} "
            "Max Norm:{maxnorm} Rate:{rate}".format(self.n_in, self.n_out, self.loss, **self.reg))

    def TestVersion(self, inpt):
        return HingeLayer(inpt, (self.w, self.b))
