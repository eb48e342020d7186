"""ConvLayer and PoolLayer (reference: theanet/layer/convpool.py:14-127)."""
import math

from .layer import Layer, Out, activation_by_name
from .weights import init_wb


class ConvLayer(Layer):
    """True convolution (filter flipped), 'valid' or 'same', bias + activation.  mode='full' is
    rejected: the reference's out_sz for it (in+f+1, convpool.py:64) is wrong.  stride > 1 runs as
    the stride-1 convolution sampled every `stride` pixels, in 'valid' mode (the reference asserts
    stride 1 for 'same', :58) and only where its bookkeeping `out_sz //= stride` (:70) agrees with
    what conv2d(subsample=...) returns, i.e. when in_sz - filter_sz + 1 divides by stride."""

    def __init__(self, inpt, wts, rand_gen, batch_sz, num_prev_maps, in_sz, num_maps, filter_sz,
                 stride, mode='valid', actvn='relu50', reg=()):
        assert (wts is not None or rand_gen is not None)
        assert mode in ("valid", "full", "same")
        if mode == 'full':
            raise NotImplementedError("ConvLayer mode='full' is unsupported (the reference's "
                                      "out_sz for it is inconsistent, convpool.py:64)")
        if mode == 'same':
            assert stride == 1, "For Same mode stride should be 1"
        if stride != 1 and (in_sz - filter_sz + 1) % stride:
            raise NotImplementedError(
                "ConvLayer stride {}: in_sz - filter_sz + 1 = {} is not a multiple of it, the "
                "reference's out_sz (convpool.py:70) would disagree with the convolution's".format(
                    stride, in_sz - filter_sz + 1))
        filter_shape = (num_maps, num_prev_maps, filter_sz, filter_sz)
        fan_in = num_prev_maps * filter_sz * filter_sz
        fan_out = num_maps * filter_sz * filter_sz
        self.W, self.b = init_wb(wts, rand_gen, filter_shape, (filter_shape[0],), fan_in, fan_out,
                                 actvn, 'Conv')
        if mode == 'same':
            shift = (filter_sz - 1) // 2            # 'full' cropped by shift (convpool.py:57-61)
            self.pad_lo = filter_sz - 1 - shift
            self.out_sz = in_sz
        else:
            self.pad_lo = 0
            self.out_sz = in_sz - filter_sz + 1
        self.full_sz = self.out_sz              # output side of the stride-1 convolution
        self.stride = stride
        self.out_sz //= stride
        self.actvn = actvn
        self.act = activation_by_name(actvn)
        self.params = [self.W, self.b]
        self.inpt = inpt
        self.num_prev_maps = num_prev_maps
        self.in_sz = in_sz
        self.num_maps = num_maps
        self.filter_sz = filter_sz
        self.mode = mode
        self.n_out = num_maps * self.out_sz ** 2
        self.reg = {"L1": 0, "L2": 0, "momentum": .95, "rate": 1, "maxnorm": 0, }
        self.reg.update(reg)
        self.args = (batch_sz, num_prev_maps, in_sz, num_maps, filter_sz, stride, mode, actvn, reg)
        self.output = Out(self, (num_maps, self.out_sz, self.out_sz))
        self.representation = (
            "Conv Maps:{:2d} Filter:{} Stride:{} Mode:{} Output:{:2d} Act:{}\n\t  L1:{L1} L2:{L2} "
            "Momentum:{momentum} Rate:{rate} Max Norm:{maxnorm}".format(
                num_maps, filter_sz, stride, mode, self.out_sz, actvn, **self.reg))

    def TestVersion(self, inpt):
        return ConvLayer(inpt, (self.W, self.b), None, *self.args)


class PoolLayer(Layer):
    def __init__(self, inpt, num_maps, in_sz, pool_sz, ignore_border=False):
        if ignore_border:
            self.out_sz = in_sz // pool_sz
        else:
            self.out_sz = math.ceil(in_sz / pool_sz)
        self.params = []
        self.inpt = inpt
        self.num_maps = num_maps
        self.in_sz = in_sz
        self.pool_sz = pool_sz
        self.ignore_border = ignore_border
        self.args = (num_maps, in_sz, pool_sz, ignore_border)
        self.n_out = num_maps * self.out_sz ** 2
        self.output = Out(self, (num_maps, self.out_sz, self.out_sz))
        self.representation = "Pool Maps:{:2d} Pool_sz:{} Border:{} Output:{:2d}".format(
            num_maps, pool_sz, "Ignore" if ignore_border else "Keep", self.out_sz)

    def TestVersion(self, inpt):
        return PoolLayer(inpt, *self.args)


class MeanLayer(Layer):
    """Global average over each map (reference: theanet/layer/convpool.py:129-144):
    (B, maps, S, S) -> (B, maps)."""

    def __init__(self, inpt, num_maps, in_sz):
        self.params = []
        self.inpt = inpt
        self.num_maps = num_maps
        self.in_sz = in_sz
        self.out_sz = 1
        self.n_out = num_maps
        self.output = Out(self, (num_maps,))
        self.representation = "Mean Maps:{:2d} Output:{:2d}".format(num_maps, self.out_sz)

    def TestVersion(self, inpt):
        return MeanLayer(inpt, self.num_maps, self.in_sz)
