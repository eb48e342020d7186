"""ctypes binding of libtheanet_b200.so (the C ABI in include/theanet_b200.h).

There is no CPU fallback: if the shared library is missing the import of this module fails, and
every compute call raises if the kernel launch fails.  torch tensors are used only as device
memory containers -- what crosses the boundary is ``tensor.data_ptr()`` and the raw stream.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libtheanet_b200.so')

# activation / rng enums (include/theanet_b200.h)
ACT_LINEAR, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SCALED_TANH, ACT_SIGMOID, ACT_SOFTPLUS = range(7)
OUT_SOFTMAX, OUT_EXPLOSS, OUT_HINGE = range(3)                      # TN_OUT_*
LOSS_NLL, LOSS_NLLSQ, LOSS_NLLTRUNC, LOSS_EXP, LOSS_HINGE = range(5)   # TN_LOSS_*
CTL_STEP, CTL_SAMPLE0, CTL_ROW0, CTL_LR_BITS, CTL_WORDS = 0, 1, 2, 3, 8
RNG_DROPOUT, RNG_FLIP, RNG_NOISE, RNG_SCALARS = range(4)


class ElasticPrm(C.Structure):
    _fields_ = [('h', C.c_int), ('sigma', C.c_int), ('translation', C.c_float),
                ('magnitude', C.c_float), ('log_zoom', C.c_float), ('angle_rad', C.c_float),
                ('zoom_on', C.c_int), ('nearest', C.c_int), ('clip_hi', C.c_double),
                ('step_offset', C.c_int), ('reserved', C.c_int)]


class ParamSeg(C.Structure):
    _fields_ = [('offset', C.c_int64), ('size', C.c_int64), ('ndim', C.c_int32),
                ('rows', C.c_int32), ('cols', C.c_int32), ('momentum', C.c_float),
                ('rate', C.c_float), ('maxnorm', C.c_float), ('l1', C.c_float), ('l2', C.c_float)]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_D = C.c_double
_U64 = C.c_uint64
_I64 = C.c_int64

# name -> (restype, argtypes); every name here must be declared in include/theanet_b200.h
SIGNATURES = {
    'tn_version': (_I, []),
    'tn_last_error': (C.c_char_p, []),
    'tn_launch_count': (C.c_uint64, []),
    'tn_device_check': (_I, [_I]),
    'tn_set_ctl': (_I, [_P, _I, _I, _I, _I, _P]),
    'tn_philox_words': (_I, [_P, _I, _I, _U64, _I, _I, _I, _P]),
    'tn_elastic_noise': (_I, [_P, _I, _U64, _P, _P]),
    'tn_elastic_field': (_I, [C.POINTER(ElasticPrm), _P, _P, _P, _U64, _P, _P, _P, _P, _P, _P]),
    'tn_elastic_warp': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _D, _P, _U64, _P, _P]),
    'tn_conv2d_fprop': (_I, [_P, _P, _P, _P] + [_I] * 9 + [_P]),
    'tn_conv2d_dgrad': (_I, [_P, _P, _P, _P] + [_I] * 9 + [_P]),
    'tn_conv2d_wgrad_workspace_bytes': (C.c_size_t, [_I] * 5),
    'tn_conv2d_wgrad': (_I, [_P, _P, _P, _P, _P] + [_I] * 7 + [_P]),
    'tn_convpool_fprop': (_I, [_P] * 5 + [_I] * 11 + [_P]),
    'tn_convpool_bwd_weights_workspace_bytes': (C.c_size_t, [_I] * 4),
    'tn_convpool_bwd_weights': (_I, [_P] * 7 + [_I] * 11 + [_P]),
    'tn_convpool_bwd_data': (_I, [_P] * 6 + [_I] * 13 + [_P]),
    'tn_convpool_small_supported': (_I, [_I] * 9),
    'tn_convpool_bwd_workspace_bytes': (C.c_size_t, [_I] * 11),
    'tn_convpool_fprop_train': (_I, [_P] * 6 + [_I] * 11 + [_P]),
    'tn_convpool_bwd': (_I, [_P] * 11 + [_I] * 13 + [_P]),
    'tn_convpool_debug_timestamps': (_I, [_P]),
    'tn_conv2d_tc_supported': (_I, [_I] * 5),
    'tn_nchw_f32_to_nhwc_bf16': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'tn_nhwc_bf16_to_nchw_f32': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'tn_conv2d_tc_pack_weights': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'tn_conv2d_tc_fprop': (_I, [_P] * 5 + [_I] * 9 + [_P]),
    'tn_conv2d_tc_dgrad': (_I, [_P] * 3 + [_I] * 7 + [_P]),
    'tn_im2col_bf16': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'tn_conv2d_tc_pack_weights_im2col': (_I, [_P, _P, _I, _I, _I, _P]),
    'tn_conv2d_tc_unpack_wgrad_im2col': (_I, [_P, _P, _I, _I, _I, _P]),
    'tn_maxpool2_nhwc_bf16': (_I, [_P, _P, _I, _I, _I, _P]),
    'tn_poolbwd_nhwc_bf16': (_I, [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    'tn_conv2d_tc_wgrad_workspace_bytes': (C.c_size_t, [_I] * 5),
    'tn_conv2d_tc_wgrad': (_I, [_P] * 5 + [_I] * 7 + [_P]),
    'tn_maxpool_fwd': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'tn_maxpool_bwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tn_dense_fwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _D, _U64, _P, _P, _F, _P]),
    'tn_dense_bwd_data': (_I, [_P, _P, _P, _I, _I, _I, _P, _I, _I, _D, _U64, _P, _P, _P]),
    'tn_dense_bwd_weights': (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    'tn_dense_bwd_data_sm': (_I, [_P, _P, _P, _I, _I, _I, _P, _I, _I, _D, _U64, _P, _P, _I, _P]),
    'tn_dense_bwd_weights_sm': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    'tn_set_dense_mode': (_I, [_I]),
    'tn_dense_debug_timestamps': (_I, [_P]),
    'tn_subsample2d': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'tn_upsample2d_zero': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'tn_meanpool_fwd': (_I, [_P, _P, _I, _I, _P]),
    'tn_meanpool_bwd': (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    'tn_color_jitter': (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _U64, _P, _P, _P]),
    'tn_aux_location_mix': (_I, [_P, _P, _I, _F, _I, _U64, _P, _P, _P]),
    'tn_concat_cols': (_I, [_P, _I, _P, _I, _P, _I, _P]),
    'tn_slice_cols': (_I, [_P, _I, _I, _I, _P, _I, _P]),
    'tn_add_inplace': (_I, [_P, _P, _I64, _P]),
    'tn_dropout_apply': (_I, [_P, _P, _I, _I, _D, _U64, _P, _P, _F, _P]),
    'tn_act_bwd': (_I, [_P, _P, _P, _I64, _I, _I, _P]),
    'tn_softmax_nll_fwd_bwd': (_I, [_P, _P, _P, _P, _I, _I, _F, _P, _P, _P, _P]),
    'tn_softmax_head_supported': (_I, [_I, _I]),
    'tn_softmax_head_fwd_bwd': (_I, [_P] * 6 + [_I, _I, _I, _F] + [_P] * 4 + [_I, _I, _I, _D, _U64,
                                                                            _P, _P]),
    'tn_softmax_head_workspace_bytes': (C.c_size_t, [_I, _I, _I]),
    'tn_softmax_head_bwd_weights': (_I, [_P] * 5 + [_I, _I, _I, _P, _P, _P]),
    'tn_softmax_test_stats': (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    'tn_output_loss_fwd_bwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P]),
    'tn_output_test_stats': (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    'tn_update_workspace_bytes': (C.c_size_t, [_I, _I64]),
    'tn_sgd_momentum_maxnorm_update': (_I, [_P, _P, _P, C.POINTER(ParamSeg), _I, _I64, _P, _F,
                                            _P, _F, _P, _P, _P]),
    'tn_allreduce_sgd_update': (_I, [_P, _P, _P, _P, _I, _I, C.POINTER(ParamSeg), _I, _I64, _I64, _P, _F, _F,
                                     _P, _P, _P]),
    'tn_peer_allreduce': (_I, [_P, _P, _I, _I, _I64, _I64, _P]),
    'tn_peer_alloc': (_I, [C.c_size_t, C.POINTER(C.c_void_p)]),
    'tn_peer_free': (_I, [_P]),
    'tn_ipc_get_handle': (_I, [_P, _P]),
    'tn_ipc_open_handle': (_I, [_P, C.POINTER(C.c_void_p)]),
    'tn_ipc_close_handle': (_I, [_P]),
    'tn_reduce_rowloss': (_I, [_P, _I, _P, _P]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "theanet_b200: {} is missing -- build it with `python -m theanet_b200.build` "
        "(there is no CPU fallback)".format(LIB_PATH))

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class TheanetB200Error(RuntimeError):
    pass


def last_error():
    return lib.tn_last_error().decode('utf-8', 'replace')


def check(rc, what=''):
    if rc != 0:
        raise TheanetB200Error('{} failed (code {}): {}'.format(what or 'tn call', rc, last_error()))


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or NULL for None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def call(name, *args):
    check(getattr(lib, name)(*args), name)


def act_code(name):
    """theanet activation name (theanet/layer/layer.py:27-54) -> (TN_ACT_*, NN)."""
    if name == 'linear':
        return ACT_LINEAR, 0
    if name == 'relu':
        return ACT_RELU, 0
    if name == 'tanh':
        return ACT_TANH, 0
    if name == 'scaled_tanh':
        return ACT_SCALED_TANH, 0
    if name == 'sigmoid':
        return ACT_SIGMOID, 0
    if name == 'softplus':
        return ACT_SOFTPLUS, 0
    if len(name) == 6 and name.startswith('relu') and name[4:].isdigit():
        return ACT_LEAKY, int(name[4:])
    raise NotImplementedError("Unknown Activation Specified: " + name)
