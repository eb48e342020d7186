"""The C-ABI boundary: libtheanet_b200.so loads, exports every entry point that
include/theanet_b200.h declares, and the ctypes table binds exactly that set (no compute calls --
this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'theanet_b200.h')) as f:
        src = f.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tn_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for need in ('tn_conv2d_fprop', 'tn_conv2d_dgrad', 'tn_conv2d_wgrad', 'tn_maxpool_fwd',
                 'tn_maxpool_bwd', 'tn_dense_fwd', 'tn_elastic_warp', 'tn_softmax_nll_fwd_bwd',
                 'tn_sgd_momentum_maxnorm_update', 'tn_convpool_fprop', 'tn_softmax_head_fwd_bwd'):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from theanet_b200 import build
    lib = ctypes.CDLL(build.build())
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.tn_version() >= 100


def test_ctypes_table_matches_header():
    from theanet_b200 import _C
    assert sorted(_C.SIGNATURES) == declared_symbols()
    assert _C.last_error() == '' or isinstance(_C.last_error(), str)


def test_argument_errors_are_reported_without_a_gpu():
    from theanet_b200 import _C
    rc = _C.lib.tn_maxpool_fwd(None, None, 1, 4, 2, 2, None)
    assert rc == -1 and 'null' in _C.last_error()
    rc = _C.lib.tn_set_dense_mode(9)
    assert rc == -1
    assert _C.lib.tn_softmax_head_supported(500, 10) == 1
    assert _C.lib.tn_softmax_head_supported(1000, 457) == 0
