"""tcgen05 implicit-GEMM convolution (conv_tc.cu) against the oracle on bf16-rounded operands.

Tolerances: operands are rounded to bfloat16 on both sides, accumulation is float32, so the only
differences are the accumulation order (~1e-6) and the bf16 rounding of the OUTPUT tensors a / dx
(2^-9 relative per element): |got - want| <= 2^-8 |want| + 1e-3 max|want|.  dW / db stay float32:
1e-3 relative.  The fused 2x2 pool is bit-exact on the rounded activations."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import theanet_oracle as O   # noqa: E402

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope='module')
def C():
    from theanet_b200 import _C
    _C.check(_C.lib.tn_device_check(0), 'tn_device_check')
    return _C


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def to_nhwc_bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a.transpose(0, 2, 3, 1))).to('cuda').to(torch.bfloat16).contiguous()


def from_nhwc(t):
    return t.to(torch.float32).cpu().numpy().transpose(0, 3, 1, 2)


def close_bf16(got, want):
    tol = np.abs(want) * 2.0 ** -8 + 1e-3 * np.max(np.abs(want))
    bad = np.abs(got - want) > tol
    assert not bad.any(), (int(bad.sum()), float(np.max(np.abs(got - want))), float(np.max(np.abs(want))))


CASES = [  # B, C, S, M, f, act, pool
    (4, 64, 16, 128, 3, 'relu05', True),     # C4 conv 2
    (5, 128, 8, 256, 3, 'relu05', True),     # C4 conv 3 (two images per tile, odd batch)
    (2, 64, 32, 64, 3, 'relu', False),       # 32-wide rows, N = 64
    (3, 128, 16, 128, 5, 'linear', True),    # 5x5 filter
    (2, 64, 4, 64, 3, 'relu10', True),       # eight 4x4 images per tile
]


@pytest.mark.parametrize('case', range(len(CASES)))
def test_conv_tc_fprop_dgrad_wgrad(C, case):
    B, Cin, S, M, f, actn, pool = CASES[case]
    rng = np.random.default_rng(500 + case)
    x = bf16_round(rng.standard_normal((B, Cin, S, S)))
    W = bf16_round(rng.standard_normal((M, Cin, f, f)) / np.sqrt(Cin * f * f))
    b = rng.standard_normal(M).astype(np.float32)
    pad_lo, out_sz = O.conv_geometry(S, f, 'same')
    assert C.lib.tn_conv2d_tc_supported(Cin, S, M, f, out_sz)
    act, nn = C.act_code(actn)
    z, cache = O.conv_forward(x.astype(np.float64), W.astype(np.float64), 'same')
    a = O.act_forward(actn, (z + b[None, :, None, None]).astype(np.float32))

    xd = to_nhwc_bf16(x)
    Wd = torch.from_numpy(W).cuda()
    bd = torch.from_numpy(b).cuda()
    Wp = torch.zeros(M * f * f * Cin, dtype=torch.bfloat16, device='cuda')
    Wpd = torch.zeros(M * f * f * Cin, dtype=torch.bfloat16, device='cuda')
    C.call('tn_conv2d_tc_pack_weights', C.ptr(Wd), C.ptr(Wp), M, Cin, f, 0, None)
    C.call('tn_conv2d_tc_pack_weights', C.ptr(Wd), C.ptr(Wpd), M, Cin, f, 1, None)
    ad = torch.zeros((B, out_sz, out_sz, M), dtype=torch.bfloat16, device='cuda')
    pd = torch.zeros((B, out_sz // 2, out_sz // 2, M), dtype=torch.bfloat16, device='cuda')
    C.call('tn_conv2d_tc_fprop', C.ptr(xd), C.ptr(Wp), C.ptr(bd), C.ptr(ad),
           C.ptr(pd) if pool else None, B, Cin, S, M, f, pad_lo, out_sz, act, nn, None)
    torch.cuda.synchronize()
    got_a = from_nhwc(ad)
    close_bf16(got_a, a)
    if pool:
        want_p, _ = O.pool_forward(got_a, 2, False)           # bit-exact on the GPU's own a
        assert np.array_equal(from_nhwc(pd), want_p)

    # dgrad / wgrad from a bf16 gradient tensor
    gz = bf16_round(rng.standard_normal(z.shape) * (rng.random(z.shape) < .5))
    dW, db, dx = O.conv_backward(gz.astype(np.float64), W.astype(np.float64), cache)
    gzd = to_nhwc_bf16(gz)
    dxd = torch.zeros((B, S, S, Cin), dtype=torch.bfloat16, device='cuda')
    C.call('tn_conv2d_tc_dgrad', C.ptr(gzd), C.ptr(Wpd), C.ptr(dxd), B, Cin, S, M, f, pad_lo, out_sz,
           None)
    torch.cuda.synchronize()
    close_bf16(from_nhwc(dxd), dx)
    nb = C.lib.tn_conv2d_tc_wgrad_workspace_bytes(B, Cin, M, f, out_sz)
    assert nb > 0
    ws = torch.zeros(nb // 4 + 4, device='cuda')
    res = []
    for _ in range(2):
        dWd, dbd = torch.zeros_like(Wd), torch.zeros_like(bd)
        C.call('tn_conv2d_tc_wgrad', C.ptr(xd), C.ptr(gzd), C.ptr(dWd), C.ptr(dbd), C.ptr(ws), B, Cin,
               S, M, f, pad_lo, out_sz, None)
        torch.cuda.synchronize()
        res.append((dWd.cpu().numpy(), dbd.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.max(np.abs(res[0][0] - dW)) <= 1e-3 * np.max(np.abs(dW))
    assert np.max(np.abs(res[0][1] - db)) <= 1e-3 * np.max(np.abs(db))


def test_layout_round_trip_and_refusals(C):
    rng = np.random.default_rng(1)
    x = bf16_round(rng.standard_normal((3, 5, 6, 7)))
    xd = torch.from_numpy(x).cuda()
    y = torch.zeros((3, 6, 7, 5), dtype=torch.bfloat16, device='cuda')
    C.call('tn_nchw_f32_to_nhwc_bf16', C.ptr(xd), C.ptr(y), 3, 5, 6, 7, None)
    back = torch.zeros_like(xd)
    C.call('tn_nhwc_bf16_to_nchw_f32', C.ptr(y), C.ptr(back), 3, 5, 6, 7, None)
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy(), x)
    assert np.array_equal(y.to(torch.float32).cpu().numpy(), x.transpose(0, 2, 3, 1))
    assert not C.lib.tn_conv2d_tc_supported(3, 32, 64, 3, 32)      # C = 3: CUDA-core path
    assert not C.lib.tn_conv2d_tc_supported(64, 28, 64, 3, 28)     # 28-wide rows do not tile
    rc = C.lib.tn_conv2d_tc_fprop(C.ptr(y), C.ptr(y), C.ptr(xd), C.ptr(y), None, 1, 3, 8, 64, 3, 1, 8,
                                  0, 0, None)
    assert rc == -5
