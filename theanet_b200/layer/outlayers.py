"""SoftmaxLayer and the output-layer protocol (reference: theanet/layer/outlayers.py:12-102).

Only the 'nll' loss is on the hot path this package covers; the other losses and output layers
(nllsq, nllNN, hinge, exp, CenteredOutLayer) are out of scope."""
from .hidden import HiddenLayer
from .layer import Out


class OutputLayer(object):
    def cost(self, y=None):
        if self.loss == "nll":
            return "nll"
        raise NotImplementedError("Loss : {} (only 'nll' is implemented)".format(self.loss))

    def features_and_predictions(self):
        return self.features, self.y_preds

    def sym_and_oth_err_rate(self, y=None):
        return "sym_err_rate", "second_stat"


class SoftmaxLayer(HiddenLayer, OutputLayer):
    def __init__(self, inpt, wts, rand_gen=None, n_in=None, n_out=None, reg=(), loss="nll"):
        HiddenLayer.__init__(self, inpt, wts, rand_gen, n_in, n_out, actvn='Softmax', reg=reg,
                             pdrop=0)
        # handles resolved by the engine: log-probabilities, probabilities, argmax
        self.logprob = Out(self, (self.n_out,))
        self.features = self.logprob            # outlayers.py:93
        self.probs = self.output
        self.y_preds = Out(self, ())
        self.kind = 'SOFTMAX'
        self.loss = loss
        if loss is not None:
            self.cost()
        self.representation = (
            "Softmax In:{:3d} Out:{:3d} Loss:{}\n\t  L1:{L1} L2:{L2} Momentum:{momentum} "
            "Max Norm:{maxnorm} Rate:{rate}".format(self.n_in, self.n_out, self.loss, **self.reg))

    def TestVersion(self, inpt):
        return SoftmaxLayer(inpt, (self.w, self.b), loss=None)
