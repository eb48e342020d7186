"""Layer classes of theanet_b200 -- the host-side mirror of the reference's layer protocol
(constructor keywords = the .prms vocabulary; attributes num_maps / out_sz / n_out / params / reg /
representation; TestVersion twins).  NeuralNet resolves the class named in a .prms entry with
``getattr(theanet_b200.layer, name)``, so everything a network definition may mention is exported
here.  The objects only describe the network: all arithmetic runs in the CUDA kernels behind
include/theanet_b200.h, launched by theanet_b200.neuralnet.
"""
from .weights import Param, init_wb, borrow, is_shared_var
from .layer import Layer, Activation, activation_by_name, Out

# data layers
from .inlayers import InputLayer, ElasticLayer
from .color import ColorLayer
# feature layers
from .convpool import ConvLayer, PoolLayer, MeanLayer
from .dropout import DropOutLayer
from .hidden import HiddenLayer
from .auxiliary import AuxConcatLayer, LocationInfo
# output layers
from .outlayers import OutputLayer, SoftmaxLayer, ExpLossLayer, HingeLayer, OUT_KINDS
from .auxiliary import SoftAuxLayer

__all__ = [
    'Param', 'init_wb', 'borrow', 'is_shared_var', 'Layer', 'Activation', 'activation_by_name', 'Out',
    'InputLayer', 'ElasticLayer', 'ColorLayer', 'ConvLayer', 'PoolLayer', 'MeanLayer', 'DropOutLayer',
    'HiddenLayer', 'AuxConcatLayer', 'LocationInfo', 'OutputLayer', 'SoftmaxLayer', 'ExpLossLayer',
    'HingeLayer', 'SoftAuxLayer', 'OUT_KINDS',
]
