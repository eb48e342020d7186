from .layer import Layer, Activation, activation_by_name, Out
from .weights import Param, init_wb, borrow, is_shared_var
from .inlayers import InputLayer, ElasticLayer
from .color import ColorLayer
from .convpool import ConvLayer, PoolLayer, MeanLayer
from .dropout import DropOutLayer
from .hidden import HiddenLayer
from .outlayers import SoftmaxLayer, ExpLossLayer, HingeLayer, OutputLayer, OUT_KINDS
from .auxiliary import SoftAuxLayer, AuxConcatLayer, LocationInfo
