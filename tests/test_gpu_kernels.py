"""Per-kernel parity: every C-ABI entry point against the CPU oracle on the same seeded inputs.

Bars: bit-exact for integer / index / comparison work (Philox words, gather indices, nearest and
bilinear warps, pool forward / tie routing, argmax); for float32 arithmetic whose summation order
differs from numpy's, max|a-b| / max|b| <= 1e-3 per tensor (north_star tolerance) -- in practice
these agree to ~1e-6."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import theanet_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-3


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.fixture(scope='module')
def C():
    from theanet_b200 import _C
    assert torch.cuda.is_available()
    _C.check(_C.lib.tn_device_check(0), 'tn_device_check')
    return _C


_KEEP = []      # device tensors whose raw pointers were handed to the C ABI stay alive per test


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    t = t.cuda()
    _KEEP.append(t)
    return t


def make_ctl(C, step=0, sample0=0, row0=0, lr=0.1):
    c = np.zeros(C.CTL_WORDS, np.int32)
    c[C.CTL_STEP], c[C.CTL_SAMPLE0], c[C.CTL_ROW0] = step, sample0, row0
    c[C.CTL_LR_BITS] = np.float32(lr).view(np.int32)
    return dev(c)


def sync():
    torch.cuda.synchronize()


# --------------------------------------------------------------------------------------------
def test_philox_words_bit_exact(C):
    for n, ns, seed, purpose, step, s0 in [(500, 7, 802165, 0, 3, 40), (784, 5, 2 ** 40 + 17, 1, 0, 0),
                                           (13, 3, 1, 2, 2 ** 31 - 1, 1000)]:
        w = torch.zeros((ns, n), dtype=torch.int32, device='cuda')
        C.call('tn_philox_words', C.ptr(w), ns, n, seed, purpose, step, s0, None)
        sync()
        want = philox.random_words(seed, purpose, step, np.arange(s0, s0 + ns), n)
        assert np.array_equal(w.cpu().numpy().view(np.uint32), want)


def elastic_prm(C, h, args):
    p = C.ElasticPrm()
    p.h, p.sigma = h, int(args.get('sigma', 1))
    p.translation, p.magnitude = float(args.get('translation', 0)), float(args.get('magnitude', 0))
    zoom = args.get('zoom', 1)
    p.zoom_on = int(zoom != 1)
    p.log_zoom = float(np.float32(np.log(zoom)))
    p.angle_rad = float(np.float32(args.get('angle', 0) * np.pi / 180))
    p.nearest = int(bool(args.get('nearest', False)))
    p.clip_hi = h - 1 - .001
    return p


def test_elastic_noise_matches_oracle_stream(C):
    h, seed, step = 28, 426405, 5
    noise = torch.zeros(2 * h * h, device='cuda')
    C.call('tn_elastic_noise', C.ptr(noise), h, seed, C.ptr(make_ctl(C, step=step)), None)
    sync()
    want = philox.elastic_noise(seed, step, 2 * h * h)
    got = noise.cpu().numpy()
    # float64 log/sincos on the device may differ from libm in the last ulp before the float32
    # rounding; equality is expected for (nearly) every value
    assert np.mean(got == want) > 0.999
    assert np.allclose(got, want, rtol=3e-7, atol=0)


ELASTIC_CASES = [
    dict(translation=2, zoom=1.1, magnitude=60, sigma=15, angle=5, nearest=True),
    dict(translation=2, zoom=1.1, magnitude=20, sigma=5, angle=5, nearest=False),
    dict(translation=3, nearest=True),
    dict(zoom=1.3, nearest=False),
    dict(angle=20, nearest=True),
    dict(magnitude=30, sigma=4, nearest=False),
]


@pytest.mark.parametrize('case', range(len(ELASTIC_CASES)))
@pytest.mark.parametrize('h', [28, 33])
def test_elastic_field_and_warp_bit_exact(C, case, h):
    args = ELASTIC_CASES[case]
    rng = np.random.default_rng(100 + case)
    noise = rng.standard_normal((2, h, h)).astype(np.float32)
    u = rng.uniform(0, 1, 8).astype(np.float32)
    ty, tx, disp = O.elastic_target(h, args, noise, u)
    prm = elastic_prm(C, h, args)
    sigma = int(args.get('sigma', 1))
    filt = dev(O.gaussian_filter(sigma)) if args.get('magnitude', 0) else None
    target = torch.zeros(2 * h * h, dtype=torch.float64, device='cuda')
    tyx = torch.zeros(2 * h * h, dtype=torch.float64, device='cuda')
    gidx = torch.zeros(h * h, dtype=torch.int32, device='cuda')
    gfrac = torch.zeros(2 * h * h, device='cuda')
    C.call('tn_elastic_field', ctypes.byref(prm), C.ptr(dev(noise)), C.ptr(dev(u)), C.ptr(filt), 0,
           None, C.ptr(target), C.ptr(tyx), C.ptr(gidx), C.ptr(gfrac), None)
    sync()
    got = tyx.cpu().numpy().reshape(2, h, h)
    assert np.max(np.abs(got[0] - ty)) < 1e-9 and np.max(np.abs(got[1] - tx)) < 1e-9
    assert np.max(np.abs(target.cpu().numpy().reshape(2, h, h) - np.indices((h, h)) - disp)) < 1e-9
    if args.get('nearest'):
        want_idx = O._iround(ty) * h + O._iround(tx)
    else:
        want_idx = ty.astype(np.int32) * h + tx.astype(np.int32)
    assert np.array_equal(gidx.cpu().numpy().reshape(h, h), want_idx)
    # warp: B images, C_ maps, invert + flip noise from the Philox stream, bit-exact
    B, C_ = 5, 2
    x = rng.uniform(0, 1, (B + 3, C_, h, h)).astype(np.float32)
    x *= (x > .5)
    seed, step, s0, row0, pflip = 99, 4, 64, 2, 0.03
    out = torch.zeros((B, C_, h, h), device='cuda')
    ctl = make_ctl(C, step=step, sample0=s0, row0=row0)
    C.call('tn_elastic_warp', C.ptr(dev(x)), None, C.ptr(ctl), B, C_, h, 1,
           1 if args.get('nearest') else 2, C.ptr(gidx), C.ptr(gfrac), pflip, None, seed,
           C.ptr(out), None)
    sync()
    fm = philox.bernoulli_mask(seed, philox.PURPOSE_FLIP, step, np.arange(s0, s0 + B), C_ * h * h,
                               pflip).reshape(B, C_, h, h)
    want = O.elastic_apply(x[row0:row0 + B], dict(args, invert_image=True), ty, tx, fm)
    assert np.array_equal(out.cpu().numpy(), want)


def test_elastic_warp_identity_index_list_and_injected_flip(C):
    rng = np.random.default_rng(7)
    B, C_, h = 6, 3, 9         # C*h*h = 243: exercises the non-multiple-of-4 tail
    x = rng.uniform(0, 1, (20, C_, h, h)).astype(np.float32)
    idx = rng.integers(0, 20, B).astype(np.int32)
    fm = (rng.uniform(0, 1, (B, C_, h, h)) < .2).astype(np.float32)
    out = torch.zeros((B, C_, h, h), device='cuda')
    C.call('tn_elastic_warp', C.ptr(dev(x)), C.ptr(dev(idx)), C.ptr(make_ctl(C)), B, C_, h, 0, 0,
           None, None, 0.0, C.ptr(dev(fm)), 0, C.ptr(out), None)
    sync()
    want = O.elastic_apply(x[idx], {}, None, None, fm)
    assert np.array_equal(out.cpu().numpy(), want)


# --------------------------------------------------------------------------------------------
CONV_CASES = [  # B, C, S, M, f, mode, act
    (6, 1, 28, 4, 3, 'valid', 'relu10'),
    (6, 4, 13, 20, 3, 'valid', 'relu05'),
    (3, 3, 12, 7, 3, 'same', 'relu05'),
    (3, 2, 11, 5, 5, 'valid', 'tanh'),
    (2, 3, 9, 6, 4, 'same', 'relu'),
    (2, 5, 10, 9, 2, 'valid', 'linear'),
    (2, 1, 64, 4, 3, 'valid', 'relu10'),
    (2, 20, 31, 20, 3, 'same', 'relu05'),
]


@pytest.mark.parametrize('case', range(len(CONV_CASES)))
def test_conv_fprop_dgrad_wgrad(C, case):
    B, Cin, S, M, f, mode, actn = CONV_CASES[case]
    rng = np.random.default_rng(200 + case)
    x = rng.standard_normal((B, Cin, S, S)).astype(np.float32)
    W = (rng.standard_normal((M, Cin, f, f)) / np.sqrt(Cin * f * f)).astype(np.float32)
    b = rng.standard_normal(M).astype(np.float32)
    pad_lo, out_sz = O.conv_geometry(S, f, mode)
    z, cache = O.conv_forward(x, W, mode)
    a = O.act_forward(actn, z + b[None, :, None, None])
    act, nn = C.act_code(actn)
    xd, Wd, bd = dev(x), dev(W), dev(b)
    out = torch.zeros((B, M, out_sz, out_sz), device='cuda')
    C.call('tn_conv2d_fprop', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out), B, Cin, S, M, f, pad_lo,
           out_sz, act, nn, None)
    sync()
    assert rel(out.cpu().numpy(), a) < 1e-5
    gz = rng.standard_normal(z.shape).astype(np.float32)
    dW, db, dx = O.conv_backward(gz, W, cache)
    gzd = dev(gz)
    dxd = torch.zeros_like(xd)
    C.call('tn_conv2d_dgrad', C.ptr(gzd), C.ptr(Wd), C.ptr(dxd), None, B, Cin, S, M, f, pad_lo,
           out_sz, 0, 0, None)
    sync()
    assert rel(dxd.cpu().numpy(), dx) < 1e-5
    # fused act' of the producing layer: x plays the role of that layer's stored output
    C.call('tn_conv2d_dgrad', C.ptr(gzd), C.ptr(Wd), C.ptr(dxd), C.ptr(xd), B, Cin, S, M, f, pad_lo,
           out_sz, *C.act_code('relu07'), None)
    sync()
    want = O.act_backward('relu07', x, x, dx)      # sign(a) == sign(z) for leaky relu
    assert rel(dxd.cpu().numpy(), want) < 1e-5
    nb = C.lib.tn_conv2d_wgrad_workspace_bytes(B, Cin, S, M, f)
    ws = torch.zeros(nb // 4 + 1, device='cuda')
    dWd, dbd = torch.zeros_like(Wd), torch.zeros_like(bd)
    C.call('tn_conv2d_wgrad', C.ptr(xd), C.ptr(gzd), C.ptr(dWd), C.ptr(dbd), C.ptr(ws), B, Cin, S,
           M, f, pad_lo, out_sz, None)
    sync()
    assert rel(dWd.cpu().numpy(), dW) < 1e-5
    assert rel(dbd.cpu().numpy(), db) < 1e-5


# --------------------------------------------------------------------------------------------
FUSED_CASES = [  # B, C, S, M, f, mode, act, pool, ignore_border
    (6, 1, 28, 4, 3, 'valid', 'relu10', 2, False),      # mnist.prms conv 1
    (6, 4, 13, 20, 3, 'valid', 'relu05', 2, False),     # mnist.prms conv 2 (edge windows 1 wide)
    (3, 3, 12, 7, 3, 'same', 'relu05', 3, False),
    (3, 2, 11, 5, 5, 'valid', 'tanh', 2, True),
    (2, 5, 14, 9, 5, 'same', 'relu', 2, False),
    (700, 4, 13, 20, 3, 'valid', 'relu05', 2, False),   # more images than CTAs
    (2, 1, 64, 4, 3, 'valid', 'relu10', 2, False),      # C5 geometry
    (37, 3, 10, 6, 3, 'valid', 'relu', 2, True),        # ignore_border, even maps-of-4 padding
    (5, 2, 33, 5, 3, 'valid', 'linear', 2, False),      # odd side > 32: window overhang guard
    (1024, 4, 13, 20, 3, 'valid', 'relu05', 2, False),  # the benchmark's conv 2 at its batch size
    (130, 1, 28, 4, 3, 'valid', 'relu50', 2, False),    # conv 1, groups of images + a ragged tail
    (9, 4, 31, 20, 3, 'valid', 'relu05', 2, False),     # C5 conv 2
    (3, 8, 12, 32, 3, 'valid', 'relu05', 2, True),      # padded pixel stride (maps/4 even)
]


@pytest.mark.parametrize('case', range(len(FUSED_CASES)))
def test_convpool_fused_fwd_bwd(C, case):
    B, Cin, S, M, f, mode, actn, p, ib = FUSED_CASES[case]
    rng = np.random.default_rng(260 + case)
    # quantised inputs and weights make exact ties inside pool windows likely
    x = (rng.integers(-3, 4, (B, Cin, S, S)) / 4).astype(np.float32)
    W = (rng.integers(-2, 3, (M, Cin, f, f)) / 4).astype(np.float32)
    b = ((2 * rng.integers(-2, 3, M) + 1) / 32).astype(np.float32)   # z is never exactly 0 (A5)
    pad_lo, out_sz = O.conv_geometry(S, f, mode)
    z, cache = O.conv_forward(x, W, mode)
    a = O.act_forward(actn, z + b[None, :, None, None])
    pooled, pcache = O.pool_forward(a, p, ib)
    P = pooled.shape[-1]
    act, nn = C.act_code(actn)
    xd, Wd, bd = dev(x), dev(W), dev(b)
    ad = torch.zeros((B, M, out_sz, out_sz), device='cuda')
    pd = torch.zeros((B, M, P, P), device='cuda')
    C.call('tn_convpool_fprop', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(ad), C.ptr(pd), B, Cin, S, M,
           f, pad_lo, out_sz, act, nn, p, P, None)
    sync()
    assert rel(ad.cpu().numpy(), a) < 1e-6
    assert np.array_equal(pd.cpu().numpy(), O.pool_forward(ad.cpu().numpy(), p, ib)[0])  # bit-exact
    # backward from the oracle's own activations so that the tie pattern is identical
    ad, pd = dev(a), dev(pooled)
    dtop = rng.standard_normal(pooled.shape).astype(np.float32)
    da = O.pool_backward(dtop, pcache)
    gz = O.act_backward(actn, z + b[None, :, None, None], a, da)
    dW, db, dx = O.conv_backward(gz, W, cache)
    nb = C.lib.tn_convpool_bwd_weights_workspace_bytes(B, Cin, M, f)
    ws = torch.zeros(nb // 4 + 1, device='cuda')
    dtd = dev(dtop)
    res = []
    for _ in range(2):
        dWd, dbd = torch.zeros_like(Wd), torch.zeros_like(bd)
        C.call('tn_convpool_bwd_weights', C.ptr(xd), C.ptr(ad), C.ptr(pd), C.ptr(dtd), C.ptr(dWd),
               C.ptr(dbd), C.ptr(ws), B, Cin, S, M, f, pad_lo, out_sz, act, nn, p, P, None)
        sync()
        res.append((dWd.cpu().numpy(), dbd.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    tol = 1e-5 if B < 100 else 1e-4        # float32 sums over up to 1024 x 121 products
    assert rel(res[0][0], dW) < tol and rel(res[0][1], db) < tol
    dxd = torch.zeros_like(xd)
    C.call('tn_convpool_bwd_data', C.ptr(ad), C.ptr(pd), C.ptr(dtd), C.ptr(Wd), C.ptr(dxd), None, B,
           Cin, S, M, f, pad_lo, out_sz, act, nn, p, P, 0, 0, None)
    sync()
    assert rel(dxd.cpu().numpy(), dx) < 1e-5
    C.call('tn_convpool_bwd_data', C.ptr(ad), C.ptr(pd), C.ptr(dtd), C.ptr(Wd), C.ptr(dxd),
           C.ptr(xd), B, Cin, S, M, f, pad_lo, out_sz, act, nn, p, P, *C.act_code('relu07'), None)
    sync()
    xnz = np.where(x == 0, 1.0, x).astype(np.float32)    # off the kink (assumption A5)
    want = O.act_backward('relu07', xnz, xnz, dx)
    got = dxd.cpu().numpy()
    assert rel(got[x != 0], want[x != 0]) < 1e-5
    # one-launch backward of the second-generation path (conv_small.cu), where it applies
    geom = (Cin, S, M, f, pad_lo, out_sz, act, p, P)
    if not C.lib.tn_convpool_small_supported(*geom):
        assert not (f == 3 and mode == 'valid' and p == 2 and actn != 'tanh')
        return
    # training-time forward: tie pattern recorded, un-pooled activations optional
    tie = torch.zeros(B * M * P * P, dtype=torch.uint8, device='cuda')
    p2 = torch.zeros((B, M, P, P), device='cuda')
    C.call('tn_convpool_fprop_train', C.ptr(xd), C.ptr(Wd), C.ptr(bd), None, C.ptr(p2), C.ptr(tie),
           B, Cin, S, M, f, pad_lo, out_sz, act, nn, p, P, None)
    a2 = torch.zeros((B, M, out_sz, out_sz), device='cuda')
    C.call('tn_convpool_fprop_train', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(a2), C.ptr(p2), None,
           B, Cin, S, M, f, pad_lo, out_sz, act, nn, p, P, None)
    sync()
    a2n = a2.cpu().numpy()
    assert rel(a2n, a) < 1e-6
    pw, (xp_, o_, _, _, n_) = O.pool_forward(a2n, p, ib)
    assert np.array_equal(p2.cpu().numpy(), pw)
    hit = (xp_.reshape(B, M, n_, 2, n_, 2) == o_[:, :, :, None, :, None])      # (B,M,P,dy,P,dx)
    want_tie = (hit[:, :, :, 0, :, 0] * 1 + hit[:, :, :, 0, :, 1] * 2 + hit[:, :, :, 1, :, 0] * 4 +
                hit[:, :, :, 1, :, 1] * 8).astype(np.uint8)
    assert np.array_equal(tie.cpu().numpy().reshape(B, M, P, P), want_tie)       # bit-exact
    # the oracle's own tie pattern as a mask: the backward must not need `a` at all
    _, (xq, oq, _, _, nq) = O.pool_forward(a, p, ib)
    hq = (xq.reshape(B, M, nq, 2, nq, 2) == oq[:, :, :, None, :, None])
    tie_o = dev((hq[:, :, :, 0, :, 0] * 1 + hq[:, :, :, 0, :, 1] * 2 + hq[:, :, :, 1, :, 0] * 4 +
                 hq[:, :, :, 1, :, 1] * 8).astype(np.uint8))
    for need_dx, below, use_tie in ((1, False, False), (1, True, True), (0, False, True),
                                    (1, False, True)):
        nb = C.lib.tn_convpool_bwd_workspace_bytes(B, *geom, need_dx)
        ws2 = torch.zeros(nb // 4 + 1, device='cuda')
        res = []
        for _ in range(3):                     # the ticket counters must come back to zero
            dWd, dbd = torch.full_like(Wd, 7.0), torch.full_like(bd, 7.0)
            dxd = torch.full_like(xd, 7.0)
            C.call('tn_convpool_bwd', C.ptr(xd), None if use_tie else C.ptr(ad),
                   C.ptr(tie_o) if use_tie else None, C.ptr(pd), C.ptr(dtd), C.ptr(Wd),
                   C.ptr(dWd), C.ptr(dbd), C.ptr(dxd) if need_dx else None,
                   C.ptr(xd) if below else None, C.ptr(ws2), B, Cin, S, M, f, pad_lo, out_sz, act,
                   nn, p, P, *(C.act_code('relu07') if below else (0, 0)), None)
            sync()
            res.append((dWd.cpu().numpy(), dbd.cpu().numpy(), dxd.cpu().numpy()))
        for r in res[1:]:
            for name, u, v in zip(('dW', 'db', 'dx'), res[0], r):
                bad = np.argwhere(~((u == v) | (np.isnan(u) & np.isnan(v))))
                assert not len(bad) and not np.isnan(u).any(), \
                    '{} (need_dx {}, below {}, tie {}): {} entries differ between launches, {} NaN; first at {}: {} vs {}'.format(
                        name, need_dx, below, use_tie, len(bad), int(np.isnan(u).sum()),
                        bad[:3].tolist(), u[tuple(bad[0])] if len(bad) else None, v[tuple(bad[0])] if len(bad) else None)
        assert rel(res[0][0], dW) < tol and rel(res[0][1], db) < tol
        if need_dx and below:
            assert rel(res[0][2][x != 0], want[x != 0]) < 1e-5
        elif need_dx:
            assert rel(res[0][2], dx) < 1e-5
        else:
            assert np.all(res[0][2] == 7.0)


def test_convpool_fused_refuses_unsupported(C):
    x = torch.zeros((1, 2, 9, 9), device='cuda')
    rc = C.lib.tn_convpool_fprop(C.ptr(x), C.ptr(x), C.ptr(x), C.ptr(x), C.ptr(x), 1, 2, 9, 4, 4, 0,
                                 6, 0, 0, 2, 3, None)
    assert rc == -5 and 'filter_sz' in C.last_error()


def test_conv_wgrad_many_images_is_deterministic(C):
    B, Cin, S, M, f = 700, 4, 13, 20, 3          # more images than CTAs: grid-stride + 2 stages
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, Cin, S, S)).astype(np.float32)
    gz = rng.standard_normal((B, M, 11, 11)).astype(np.float32)
    W = np.zeros((M, Cin, f, f), np.float32)
    _, cache = O.conv_forward(x.astype(np.float64), W.astype(np.float64), 'valid')
    dW, db, _ = O.conv_backward(gz.astype(np.float64), W.astype(np.float64), cache, need_dx=False)
    nb = C.lib.tn_conv2d_wgrad_workspace_bytes(B, Cin, S, M, f)
    ws = torch.zeros(nb // 4 + 1, device='cuda')
    xd, gzd = dev(x), dev(gz)
    res = []
    for _ in range(2):
        dWd, dbd = torch.zeros((M, Cin, f, f), device='cuda'), torch.zeros(M, device='cuda')
        C.call('tn_conv2d_wgrad', C.ptr(xd), C.ptr(gzd), C.ptr(dWd), C.ptr(dbd), C.ptr(ws), B, Cin,
               S, M, f, 0, 11, None)
        sync()
        res.append((dWd.cpu().numpy(), dbd.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert rel(res[0][0], dW) < 1e-4 and rel(res[0][1], db) < 1e-4


def test_conv_wgrad_row_bands_large_image(C):
    """C4 conv 1: 3 -> 64 maps on 32x32 ('same'): dL/dz of one image (256 KB) does not fit in shared
    memory, the direct wgrad kernel walks it in bands of output rows."""
    B, Cin, S, M, f = 5, 3, 32, 64, 3
    rng = np.random.default_rng(77)
    x = rng.standard_normal((B, Cin, S, S)).astype(np.float32)
    gz = rng.standard_normal((B, M, S, S)).astype(np.float32)
    W = np.zeros((M, Cin, f, f), np.float32)
    pad_lo, out_sz = O.conv_geometry(S, f, 'same')
    _, cache = O.conv_forward(x.astype(np.float64), W.astype(np.float64), 'same')
    dW, db, _ = O.conv_backward(gz.astype(np.float64), W.astype(np.float64), cache, need_dx=False)
    nb = C.lib.tn_conv2d_wgrad_workspace_bytes(B, Cin, S, M, f)
    ws = torch.zeros(nb // 4 + 1, device='cuda')
    dWd, dbd = torch.zeros((M, Cin, f, f), device='cuda'), torch.zeros(M, device='cuda')
    C.call('tn_conv2d_wgrad', C.ptr(dev(x)), C.ptr(dev(gz)), C.ptr(dWd), C.ptr(dbd), C.ptr(ws), B, Cin,
           S, M, f, pad_lo, out_sz, None)
    sync()
    assert rel(dWd.cpu().numpy(), dW) < 1e-5 and rel(dbd.cpu().numpy(), db) < 1e-5


def test_conv_direct_path_refuses_oversized_shapes(C):
    x = torch.zeros((1, 64, 32, 32), device='cuda')
    W = torch.zeros((128, 64, 3, 3), device='cuda')
    ws = torch.zeros(16, device='cuda')
    rc = C.lib.tn_conv2d_wgrad(C.ptr(x), C.ptr(x), C.ptr(W), C.ptr(W), C.ptr(ws), 1, 64, 32, 128, 3,
                               1, 32, None)
    assert rc == -5 and 'implicit-GEMM' in C.last_error()


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize('S,p,ib', [(26, 2, False), (11, 2, False), (11, 2, True), (7, 3, False),
                                    (62, 2, False), (9, 4, True), (5, 5, False)])
def test_pool_fwd_bwd_bit_exact_with_ties(C, S, p, ib):
    rng = np.random.default_rng(S * 10 + p)
    B, C_ = 3, 4
    x = rng.integers(-3, 3, size=(B, C_, S, S)).astype(np.float32)         # many ties
    x[0, 0] = -5.0                                                         # all-negative, all-tied
    out, cache = O.pool_forward(x, p, ib)
    o = O.pool_out_size(S, p, ib)
    xd = dev(x)
    outd = torch.zeros((B, C_, o, o), device='cuda')
    C.call('tn_maxpool_fwd', C.ptr(xd), C.ptr(outd), B * C_, S, p, o, None)
    sync()
    assert np.array_equal(outd.cpu().numpy(), out)
    dout = rng.standard_normal(out.shape).astype(np.float32)
    dx = O.pool_backward(dout, cache)
    dxd = torch.zeros_like(xd)
    C.call('tn_maxpool_bwd', C.ptr(dev(dout)), C.ptr(xd), C.ptr(outd), C.ptr(dxd), B * C_, S, p, o,
           0, 0, None)
    sync()
    assert np.array_equal(dxd.cpu().numpy(), dx)
    # fused activation derivative (leaky relu, slope from the sign of the stored output)
    C.call('tn_maxpool_bwd', C.ptr(dev(dout)), C.ptr(xd), C.ptr(outd), C.ptr(dxd), B * C_, S, p, o,
           *C.act_code('relu10'), None)
    sync()
    want = O.act_backward('relu10', x, x, dx)
    assert np.array_equal(dxd.cpu().numpy(), want)


# --------------------------------------------------------------------------------------------
DENSE_CASES = [(20, 720, 500, 'relu01', .5), (128, 784, 1000, 'relu10', .5), (33, 500, 10, 'linear', 0),
               (17, 1000, 457, 'linear', 0), (64, 100, 36, 'tanh', .25), (5, 7, 3, 'sigmoid', 0)]


@pytest.mark.parametrize('case', range(len(DENSE_CASES)))
def test_dense_fwd_bwd(C, case):
    B, n_in, n_out, actn, pdrop = DENSE_CASES[case]
    rng = np.random.default_rng(300 + case)
    x = rng.standard_normal((B, n_in)).astype(np.float32)
    W = (rng.standard_normal((n_in, n_out)) / np.sqrt(n_in)).astype(np.float32)
    b = rng.standard_normal(n_out).astype(np.float32)
    seed, step, s0 = 802165, 11, 256
    ctl = make_ctl(C, step=step, sample0=s0)
    act, nn = C.act_code(actn)
    a = O.act_forward(actn, x @ W + b)
    mask = philox.bernoulli_mask(seed, philox.PURPOSE_DROPOUT, step, np.arange(s0, s0 + B), n_out,
                                 1 - pdrop) if pdrop else np.ones_like(a)
    xd, Wd, bd = dev(x), dev(W), dev(b)
    out = torch.zeros((B, n_out), device='cuda')
    C.call('tn_dense_fwd', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out), B, n_in, n_out, act, nn,
           1. - pdrop, seed, C.ptr(ctl), None, 1.0, None)
    sync()
    got = out.cpu().numpy()
    assert np.array_equal(got == 0, (a * mask) == 0) or pdrop == 0      # same mask pattern
    assert rel(got, a * mask) < 1e-5
    # test twin: no mask, scaled by (1 - pdrop)
    C.call('tn_dense_fwd', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out), B, n_in, n_out, act, nn,
           1.0, 0, C.ptr(ctl), None, float(1 - pdrop), None)
    sync()
    assert rel(out.cpu().numpy(), a * np.float32(1 - pdrop)) < 1e-5
    # injected mask
    mi = (rng.uniform(0, 1, a.shape) < .7).astype(np.float32)
    C.call('tn_dense_fwd', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out), B, n_in, n_out, act, nn,
           1.0, 0, C.ptr(ctl), C.ptr(dev(mi)), 1.0, None)
    sync()
    assert rel(out.cpu().numpy(), a * mi) < 1e-5
    # backward
    g = rng.standard_normal((B, n_out)).astype(np.float32)
    gd = dev(g)
    dW, db = torch.zeros_like(Wd), torch.zeros_like(bd)
    C.call('tn_dense_bwd_weights', C.ptr(xd), C.ptr(gd), C.ptr(dW), C.ptr(db), B, n_in, n_out, None)
    sync()
    assert rel(dW.cpu().numpy(), x.T @ g) < 1e-5
    assert rel(db.cpu().numpy(), g.sum(0)) < 1e-5
    dx = torch.zeros_like(xd)
    C.call('tn_dense_bwd_data', C.ptr(gd), C.ptr(Wd), C.ptr(dx), B, n_in, n_out, None, 0, 0, 1.0,
           0, C.ptr(ctl), None, None)
    sync()
    assert rel(dx.cpu().numpy(), g @ W.T) < 1e-5
    # fused: previous dense layer with dropout + leaky relu; prev_out = its stored masked output
    pm = philox.bernoulli_mask(77, philox.PURPOSE_DROPOUT, step, np.arange(s0, s0 + B), n_in, .5)
    prev_a = O.act_forward('relu01', x)
    prev_out = (prev_a * pm).astype(np.float32)
    C.call('tn_dense_bwd_data', C.ptr(gd), C.ptr(Wd), C.ptr(dx), B, n_in, n_out, C.ptr(dev(prev_out)),
           *C.act_code('relu01'), .5, 77, C.ptr(ctl), None, None)
    sync()
    want = O.act_backward('relu01', x, prev_a, (g @ W.T) * pm)
    assert rel(dx.cpu().numpy(), want) < 1e-5


SPLITK_CASES = [(1024, 720, 500, 'relu01', .5),      # mnist.prms hidden layer at the bench batch size
                (512, 4500, 500, 'relu01', .5),      # C5: 64x64 images, 4500 -> 500
                (1024, 784, 1000, 'relu10', .5),     # 3flat.prms
                (20, 720, 500, 'relu50', 0.),        # the shipped batch size
                (130, 36, 64, 'linear', 0.),         # one narrow tile, two k-blocks, ragged rows
                (257, 1000, 132, 'relu05', .25)]     # ragged in every dimension


@pytest.mark.parametrize('case', range(len(SPLITK_CASES)))
def test_dense_split_k_tensor_core_path(C, case):
    """The default float32 dense path: 128-wide tiles, K split over a thread-block cluster, partial
    tiles exchanged through distributed shared memory and added in split order
    (gemm_tc_sk_kernel).  Against float64 NumPy, twice (determinism), and against the
    first-generation tensor-core path (tn_set_dense_mode(4))."""
    B, n_in, n_out, actn, pdrop = SPLITK_CASES[case]
    rng = np.random.default_rng(900 + case)
    x = rng.standard_normal((B, n_in)).astype(np.float32)
    W = (rng.standard_normal((n_in, n_out)) / np.sqrt(n_in)).astype(np.float32)
    b = rng.standard_normal(n_out).astype(np.float32)
    g = rng.standard_normal((B, n_out)).astype(np.float32)
    seed, step, s0 = 4711, 5, 1024
    ctl = make_ctl(C, step=step, sample0=s0)
    act, nn = C.act_code(actn)
    x64, W64, g64 = x.astype(np.float64), W.astype(np.float64), g.astype(np.float64)
    a = O.act_forward(actn, (x64 @ W64 + b).astype(np.float32))
    mask = philox.bernoulli_mask(seed, philox.PURPOSE_DROPOUT, step, np.arange(s0, s0 + B), n_out,
                                 1 - pdrop) if pdrop else np.ones_like(a)
    pm = philox.bernoulli_mask(77, philox.PURPOSE_DROPOUT, step, np.arange(s0, s0 + B), n_in, .5)
    prev_a = O.act_forward('relu01', x)
    prev_out = (prev_a * pm).astype(np.float32)
    xd, Wd, bd, gd, pod = dev(x), dev(W), dev(b), dev(g), dev(prev_out)
    runs = []
    for _ in range(2):
        out = torch.full((B, n_out), 7.0, device='cuda')
        dx = torch.full((B, n_in), 7.0, device='cuda')
        dW, db = torch.full_like(Wd, 7.0), torch.full_like(bd, 7.0)
        C.call('tn_dense_fwd', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out), B, n_in, n_out, act, nn,
               1. - pdrop, seed, C.ptr(ctl), None, 1.0, None)
        C.call('tn_dense_bwd_data', C.ptr(gd), C.ptr(Wd), C.ptr(dx), B, n_in, n_out, C.ptr(pod),
               *C.act_code('relu01'), .5, 77, C.ptr(ctl), None, None)
        C.call('tn_dense_bwd_weights', C.ptr(xd), C.ptr(gd), C.ptr(dW), C.ptr(db), B, n_in, n_out,
               None)
        sync()
        runs.append([t.cpu().numpy() for t in (out, dx, dW, db)])
    for u, v in zip(*runs):
        assert np.array_equal(u, v)
    out, dx, dW, db = runs[0]
    assert np.array_equal(out == 0, (a * mask) == 0) or pdrop == 0
    assert rel(out, a * mask) < 1e-5
    assert rel(dx, O.act_backward('relu01', x, prev_a, ((g64 @ W64.T) * pm).astype(np.float32))) < 1e-5
    assert rel(dW, x64.T @ g64) < 1e-5 and rel(db, g64.sum(0)) < 1e-5
    # the first-generation path (per-k-block promotion) on the same inputs
    out1 = torch.zeros((B, n_out), device='cuda')
    C.call('tn_set_dense_mode', 4)
    try:
        C.call('tn_dense_fwd', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out1), B, n_in, n_out, act, nn,
               1. - pdrop, seed, C.ptr(ctl), None, 1.0, None)
        sync()
    finally:
        C.call('tn_set_dense_mode', 0)
    assert rel(out, out1.cpu().numpy()) < 1e-5


@pytest.mark.parametrize('B,n_in,n_out,pdrop', [(1024, 500, 10, .5), (20, 500, 10, .5), (33, 36, 11, 0.),
                                                (17, 1000, 32, .25), (5, 7, 3, 0.), (300, 130, 2, .5)])
def test_softmax_head_fused(C, B, n_in, n_out, pdrop):
    rng = np.random.default_rng(B + n_in)
    seed, step, s0, row0 = 991, 3, 64, 5
    zprev = rng.standard_normal((B, n_in)).astype(np.float32)
    pm = philox.bernoulli_mask(seed, philox.PURPOSE_DROPOUT, step, np.arange(s0, s0 + B), n_in,
                               1 - pdrop) if pdrop else np.ones((B, n_in), np.float32)
    prev_a = O.act_forward('relu01', zprev)
    h = (prev_a * pm).astype(np.float32)                  # stored masked output of the layer below
    W = (rng.standard_normal((n_in, n_out)) / np.sqrt(n_in)).astype(np.float32)
    b = rng.standard_normal(n_out).astype(np.float32)
    y = rng.integers(0, n_out, B + row0).astype(np.int32)
    yb = y[row0:row0 + B]
    ctl = make_ctl(C, step=step, sample0=s0, row0=row0)
    inv = 1.0 / (2 * B)
    z = h.astype(np.float64) @ W.astype(np.float64) + b
    lp = O.log_softmax(z)
    g = np.exp(lp)
    g[np.arange(B), yb] -= 1
    g *= inv
    dh = O.act_backward('relu01', zprev, prev_a, (g @ W.astype(np.float64).T) * pm)
    hd, Wd, bd, yd = dev(h), dev(W), dev(b), dev(y)
    lpd = torch.zeros((B, n_out), device='cuda')
    gd = torch.zeros((B, n_out), device='cuda')
    rl = torch.zeros(B, device='cuda')
    dhd = torch.full((B, n_in), -7.0, device='cuda')
    assert C.lib.tn_softmax_head_supported(n_in, n_out)
    C.call('tn_softmax_head_fwd_bwd', C.ptr(hd), C.ptr(Wd), C.ptr(bd), C.ptr(yd), None, C.ptr(ctl),
           B, n_in, n_out, inv, C.ptr(lpd), C.ptr(gd), C.ptr(rl), C.ptr(dhd), 1,
           *C.act_code('relu01'), 1 - pdrop, seed, None, None)
    sync()
    assert rel(lpd.cpu().numpy(), lp) < 1e-5
    assert rel(gd.cpu().numpy(), g) < 1e-5
    assert rel(rl.cpu().numpy(), -lp[np.arange(B), yb]) < 1e-5
    assert rel(dhd.cpu().numpy(), dh) < 1e-5
    # plain dh (nothing fused below) and index-list labels
    idx = rng.permutation(B + row0)[:B].astype(np.int32)
    C.call('tn_softmax_head_fwd_bwd', C.ptr(hd), C.ptr(Wd), C.ptr(bd), C.ptr(yd), C.ptr(dev(idx)),
           C.ptr(ctl), B, n_in, n_out, inv, C.ptr(lpd), C.ptr(gd), C.ptr(rl), C.ptr(dhd), 0, 0, 0,
           1.0, 0, None, None)
    sync()
    g2 = np.exp(lp)
    g2[np.arange(B), y[idx]] -= 1
    g2 *= inv
    assert rel(gd.cpu().numpy(), g2) < 1e-5
    assert rel(dhd.cpu().numpy(), g2 @ W.astype(np.float64).T) < 1e-5
    # weights: twice through the same workspace (tickets must return to zero), bit-identical
    nb = C.lib.tn_softmax_head_workspace_bytes(B, n_in, n_out)
    ws = torch.zeros(nb // 4 + 1, device='cuda')
    gd = dev(g.astype(np.float32))
    res = []
    for _ in range(2):
        dW, db = torch.zeros_like(Wd), torch.zeros_like(bd)
        nll = torch.zeros(1, device='cuda')
        C.call('tn_softmax_head_bwd_weights', C.ptr(hd), C.ptr(gd), C.ptr(dW), C.ptr(db), C.ptr(ws),
               B, n_in, n_out, C.ptr(rl), C.ptr(nll), None)
        sync()
        assert abs(float(nll) - float(rl.sum())) <= 1e-5 * abs(float(rl.sum()))
        res.append((dW.cpu().numpy(), db.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    g32 = g.astype(np.float32).astype(np.float64)
    assert rel(res[0][0], h.astype(np.float64).T @ g32) < 1e-5
    assert rel(res[0][1], g32.sum(0)) < 1e-5


def test_dropout_apply_and_act_bwd(C):
    rng = np.random.default_rng(9)
    B, n = 9, 1001
    x = rng.standard_normal((B, n)).astype(np.float32)
    seed, step, s0 = 5, 2, 10
    ctl = make_ctl(C, step=step, sample0=s0)
    out = torch.zeros((B, n), device='cuda')
    C.call('tn_dropout_apply', C.ptr(dev(x)), C.ptr(out), B, n, .8, seed, C.ptr(ctl), None, 1.0, None)
    sync()
    m = philox.bernoulli_mask(seed, philox.PURPOSE_DROPOUT, step, np.arange(s0, s0 + B), n, .8)
    assert np.array_equal(out.cpu().numpy(), x * m)
    C.call('tn_dropout_apply', C.ptr(dev(x)), C.ptr(out), B, n, 1.0, 0, C.ptr(ctl), None, 0.8, None)
    sync()
    assert np.array_equal(out.cpu().numpy(), x * np.float32(.8))
    for actn in ['relu05', 'tanh', 'sigmoid', 'scaled_tanh', 'softplus', 'relu', 'linear']:
        z = rng.standard_normal((B, n)).astype(np.float32)
        a = O.act_forward(actn, z)
        g = rng.standard_normal((B, n)).astype(np.float32)
        gz = torch.zeros((B, n), device='cuda')
        C.call('tn_act_bwd', C.ptr(dev(g)), C.ptr(dev(a)), C.ptr(gz), B * n, *C.act_code(actn), None)
        sync()
        assert rel(gz.cpu().numpy(), O.act_backward(actn, z, a, g)) < 1e-5


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,n', [(20, 10), (128, 457), (7, 3), (1024, 10)])
def test_softmax_nll_and_test_stats(C, B, n):
    rng = np.random.default_rng(B + n)
    z = (3 * rng.standard_normal((B, n))).astype(np.float32)
    z[0, :] = 1.5                                   # full tie: argmax must be the first index
    ycorp = rng.integers(0, n, B + 5).astype(np.int32)
    ycorp[5], ycorp[6] = 0, n - 1
    row0 = 5
    ctl = make_ctl(C, row0=row0)
    y = ycorp[row0:row0 + B]
    lp = O.log_softmax(z)
    g = np.exp(lp)
    g[np.arange(B), y] -= 1
    g /= np.float32(2 * B)
    zd, yd = dev(z), dev(ycorp)
    lpd, gd = torch.zeros((B, n), device='cuda'), torch.zeros((B, n), device='cuda')
    rl = torch.zeros(B, device='cuda')
    C.call('tn_softmax_nll_fwd_bwd', C.ptr(zd), C.ptr(yd), None, C.ptr(ctl), B, n, 1. / (2 * B),
           C.ptr(lpd), C.ptr(gd), C.ptr(rl), None)
    nll = torch.zeros(1, device='cuda')
    C.call('tn_reduce_rowloss', C.ptr(rl), B, C.ptr(nll), None)
    sync()
    assert rel(lpd.cpu().numpy(), lp) < 1e-5
    assert rel(gd.cpu().numpy(), g) < 1e-5
    assert rel(rl.cpu().numpy(), -lp[np.arange(B), y]) < 1e-5
    assert abs(nll.item() - (-lp[np.arange(B), y]).sum()) < 1e-4 * B
    preds = torch.zeros(B, dtype=torch.int64, device='cuda')
    stats = torch.zeros(2 + 2 * B, device='cuda')
    idx = np.arange(row0, row0 + B).astype(np.int32)            # same rows through the index path
    C.call('tn_softmax_test_stats', C.ptr(zd), C.ptr(yd), C.ptr(dev(idx)), None, B, n, C.ptr(lpd),
           C.ptr(preds), C.ptr(stats), None)
    sync()
    want_pred = np.argmax(z, axis=1)
    assert np.array_equal(preds.cpu().numpy(), want_pred)
    st = stats.cpu().numpy()
    assert abs(st[0] - np.mean(want_pred != y)) < 1e-6
    assert abs(st[1] - np.mean(np.exp(lp)[np.arange(B), y])) < 1e-5


# --------------------------------------------------------------------------------------------
def test_update_all_ranks_l1_l2_frozen(C):
    rng = np.random.default_rng(11)
    shapes = [((20, 4, 3, 3), dict(maxnorm=.9)), ((20,), dict(maxnorm=.05)),
              ((37, 50), dict(maxnorm=1.1, L2=.001)), ((50,), dict(L1=.01)),
              ((50, 11), dict(rate=0, L2=.1)), ((11,), dict(rate=.5, momentum=.8)),
              ((6, 4), dict(maxnorm=2.))]
    offs, total = [], 0
    for shp, _ in shapes:
        offs.append(total)
        total += (int(np.prod(shp)) + 3) // 4 * 4
    theta, vel, grad = [rng.standard_normal(total).astype(np.float32) for _ in range(3)]
    vel *= 5
    segs = (C.ParamSeg * len(shapes))()
    lr = np.float32(.1)
    want_t, want_v = theta.copy(), vel.copy()
    wtcost = 0.
    for i, (shp, r) in enumerate(shapes):
        reg = dict(O.DEFAULT_REG, **r)
        n = int(np.prod(shp))
        o = offs[i]
        th = theta[o:o + n].reshape(shp).copy()
        if shp == (6, 4):
            th[:, 1] = 0                       # zero-norm column: scale must be exactly 1
            theta[o:o + n] = th.ravel()
            vel[o:o + n].reshape(shp)[:, 1] = 0
            want_t[o:o + n] = th.ravel()
            want_v[o:o + n] = vel[o:o + n]
        if shp == (50,):
            th[:5] = 0                         # L1 at theta == 0 contributes sgn = 0
            theta[o:o + n] = th.ravel()
            want_t[o:o + n] = th.ravel()
        s = segs[i]
        s.offset, s.size, s.ndim = o, n, len(shp)
        s.rows, s.cols = (shp[0], n // shp[0]) if len(shp) > 1 else (1, n)
        s.momentum, s.rate, s.maxnorm = reg['momentum'], reg['rate'], reg['maxnorm']
        s.l1, s.l2 = reg['L1'], reg['L2']
        wtcost += reg['L1'] * np.abs(th).sum() + reg['L2'] * (th * th).sum()
        if reg['rate']:
            gk = grad[o:o + n].reshape(shp)
            if reg['L1']:
                gk = gk + np.float32(reg['L1']) * np.sign(th)
            if reg['L2']:
                gk = gk + np.float32(2 * reg['L2']) * th
            t2, v2 = O.sgd_update(th, vel[o:o + n].reshape(shp), gk, reg, lr)
            want_t[o:o + n], want_v[o:o + n] = t2.ravel(), v2.ravel()
    td, vd, gd = dev(theta), dev(vel), dev(grad)
    nb = C.lib.tn_update_workspace_bytes(len(shapes), total)
    ws = torch.zeros(nb // 4 + 1, device='cuda')
    nll = dev(np.array([12.5], np.float32))
    cost = torch.zeros(1, device='cuda')
    C.call('tn_sgd_momentum_maxnorm_update', C.ptr(td), C.ptr(vd), C.ptr(gd), segs, len(shapes),
           total, C.ptr(make_ctl(C, lr=lr)), 1.0, C.ptr(nll), 0.25, C.ptr(cost), C.ptr(ws), None)
    sync()
    assert rel(td.cpu().numpy(), want_t) < 1e-6
    assert rel(vd.cpu().numpy(), want_v) < 1e-6
    assert abs(cost.item() - (12.5 * .25 + wtcost)) < 1e-4 * (1 + abs(wtcost))
    # frozen segment untouched, bit for bit
    o = offs[4]
    assert np.array_equal(td.cpu().numpy()[o:o + 550], theta[o:o + 550])
    assert np.array_equal(vd.cpu().numpy()[o:o + 550], vel[o:o + 550])


OUT_CASES = [('SoftmaxLayer', 'nllsq'), ('SoftmaxLayer', 'nll20'), ('SoftmaxLayer', 'nll00'),
             ('SoftmaxLayer', 'nllxx'), ('SoftmaxLayer', 'nll'), ('ExpLossLayer', 'exp'),
             ('HingeLayer', 'hinge')]


@pytest.mark.parametrize('kind,loss', OUT_CASES)
@pytest.mark.parametrize('B,n', [(7, 5), (64, 10), (33, 457)])
def test_output_layers_and_losses(C, kind, loss, B, n):
    """tn_output_loss_fwd_bwd / tn_output_test_stats against the oracle's output_views /
    output_loss (outlayers.py:38-64,66-80,105-147), which tests/test_golden_ref.py pins to the
    reference's own graph.  'nll00' = threshold 0 (log = -inf: no row contributes), 'nllxx' =
    unreadable threshold = plain NLL (outlayers.py:20-27)."""
    from theanet_b200.layer.outlayers import OUT_KINDS, OutputLayer
    rng = np.random.default_rng(B * n + len(loss))
    z = (1.5 * rng.standard_normal((B, n))).astype(np.float32)
    z[0, :] = 0.25                                  # full tie: argmax = first index, hinge margins all 1
    ycorp = rng.integers(0, n, B + 5).astype(np.int32)
    ycorp[5], ycorp[6] = 0, n - 1
    row0 = 5
    ctl = make_ctl(C, row0=row0)
    y = ycorp[row0:row0 + B].astype(np.int64)
    feat, lp, probs = O.output_views(kind, z)
    Bg = 2 * B
    cost, g = O.output_loss(kind, loss, z, lp, y, Bg)
    lyr = OutputLayer()
    lyr.loss = loss
    code, log_thr = lyr.cost()
    k = OUT_KINDS[{'SoftmaxLayer': 'SOFTMAX', 'ExpLossLayer': 'ExpLoss', 'HingeLayer': 'Hinge'}[kind]]
    zd, yd = dev(z), dev(ycorp)
    fd, lpd, gd = (torch.full((B, n), -9., device='cuda') for _ in range(3))
    rl = torch.zeros(B, device='cuda')
    C.call('tn_output_loss_fwd_bwd', C.ptr(zd), C.ptr(yd), None, C.ptr(ctl), B, n, k, code, log_thr,
           1. / Bg, C.ptr(fd), C.ptr(lpd), C.ptr(gd), C.ptr(rl), None)
    sync()
    assert rel(fd.cpu().numpy(), feat) < 1e-5
    assert rel(lpd.cpu().numpy(), lp) < 1e-5
    assert rel(gd.cpu().numpy(), g) < 1e-5 or (np.abs(g).max() == 0 and gd.abs().max().item() == 0)
    assert abs(rl.sum().item() / Bg - cost) <= 1e-5 * max(abs(cost), 1e-3)
    # which rows are active (hinge margins / truncation) is comparison work: exact
    assert np.array_equal(gd.cpu().numpy() == 0, g == 0)
    # test twin through the index-list path
    preds = torch.zeros(B, dtype=torch.int64, device='cuda')
    stats = torch.zeros(2 + 2 * B, device='cuda')
    idx = np.arange(row0, row0 + B).astype(np.int32)
    fd.fill_(-9.)
    lpd.fill_(-9.)
    C.call('tn_output_test_stats', C.ptr(zd), C.ptr(yd), C.ptr(dev(idx)), None, B, n, k, C.ptr(fd),
           C.ptr(lpd), C.ptr(preds), C.ptr(stats), None)
    sync()
    want_pred = np.argmax(z, axis=1)
    assert np.array_equal(preds.cpu().numpy(), want_pred)
    assert rel(fd.cpu().numpy(), feat) < 1e-5 and rel(lpd.cpu().numpy(), lp) < 1e-5
    st = stats[:2].cpu().numpy()
    assert abs(st[0] - np.mean(want_pred != y)) < 1e-6
    assert abs(st[1] - np.mean(probs[np.arange(B), y])) < 1e-5 * max(1., abs(np.mean(probs[np.arange(B), y])))


def test_output_loss_rejects_mismatched_kind_and_loss(C):
    z = dev(np.zeros((4, 3), np.float32))
    y = dev(np.zeros(4, np.int32))
    ctl = make_ctl(C)
    o = torch.zeros((4, 3), device='cuda')
    rl = torch.zeros(4, device='cuda')
    rc = C.lib.tn_output_loss_fwd_bwd(C.ptr(z), C.ptr(y), None, C.ptr(ctl), 4, 3, C.OUT_HINGE,
                                      C.LOSS_NLL, 0.0, 1.0, C.ptr(o), None, C.ptr(o), C.ptr(rl), None)
    assert rc == -5 and b'does not take' in C.lib.tn_last_error()      # TN_ERR_UNSUPPORTED


@pytest.mark.parametrize('planes,S,act', [(12, 8, 'relu10'), (300, 13, 'tanh'), (5, 1, 'linear'),
                                          (64, 32, 'relu')])
def test_meanpool_fwd_bwd(C, planes, S, act):
    """MeanLayer (convpool.py:129-144): mean over each map; backward spreads g / S^2 and applies
    the activation derivative of the layer below."""
    rng = np.random.default_rng(planes + S)
    z = rng.standard_normal((planes, S, S)).astype(np.float32)
    a = O.act_forward(act, z)
    g = rng.standard_normal(planes).astype(np.float32)
    want = a.sum(axis=(1, 2), dtype=np.float32) / np.float32(S * S)
    want_dx = O.act_backward(act, z, a, np.broadcast_to((g / np.float32(S * S))[:, None, None], a.shape))
    ad, gd = dev(a), dev(g)
    out = torch.zeros(planes, device='cuda')
    dx = torch.full((planes, S, S), -3., device='cuda')
    C.call('tn_meanpool_fwd', C.ptr(ad), C.ptr(out), planes, S, None)
    C.call('tn_meanpool_bwd', C.ptr(gd), C.ptr(ad), C.ptr(dx), planes, S, *C.act_code(act), None)
    sync()
    assert rel(out.cpu().numpy(), want) < 1e-6
    assert rel(dx.cpu().numpy(), want_dx) < 1e-6
    C.call('tn_meanpool_bwd', C.ptr(gd), None, C.ptr(dx), planes, S, 0, 0, None)   # nothing fused
    sync()
    assert np.array_equal(dx.cpu().numpy(),
                          np.broadcast_to((g / np.float32(S * S))[:, None, None], a.shape))


@pytest.mark.parametrize('B,Cm,S,balance,gamma,maxval', [(6, 3, 8, 1.5, 1.4, 1.), (33, 1, 28, 1., 2., 255.),
                                                         (4, 5, 64, 2., 1., 2.)])
def test_color_jitter(C, B, Cm, S, balance, gamma, maxval):
    """ColorLayer (color.py:9-52) on its Philox stream and with injected draws, against the
    oracle's color_jitter / philox.color_uniforms."""
    rng = np.random.default_rng(B + S)
    x = (rng.random((B, Cm, S, S)) * maxval).astype(np.float32)
    x[0, 0, 0, :4] = [0., maxval, 2 * maxval, -1.]             # clip on both sides, exact 0 and 1
    seed, step, s0 = 4711, 9, 40
    ctl = make_ctl(C, step=step, sample0=s0)
    prm = {'balance': balance, 'gamma': gamma, 'maxval': maxval}
    u = philox.color_uniforms(seed, step, np.arange(s0, s0 + B), Cm)
    assert u.shape == (B, Cm, 3) and np.all(np.abs(u) < 1)
    want = O.color_jitter(x, prm, u)
    lb, lg = float(np.float32(np.log(balance))), float(np.float32(np.log(gamma)))
    xd = dev(x)
    out = torch.full_like(xd, -5.)
    C.call('tn_color_jitter', C.ptr(xd), C.ptr(out), B, Cm, S, lb, lg, maxval, seed, C.ptr(ctl), None, None)
    sync()
    assert rel(out.cpu().numpy(), want) < 1e-5
    u2 = rng.uniform(-1, 1, (B, Cm, 3)).astype(np.float32)
    C.call('tn_color_jitter', C.ptr(xd), C.ptr(xd), B, Cm, S, lb, lg, maxval, 0, None, C.ptr(dev(u2)), None)
    sync()                                                         # in place, injected draws
    assert rel(xd.cpu().numpy(), O.color_jitter(x, prm, u2)) < 1e-5


def test_aux_location_mix_and_column_plumbing(C):
    """auxiliary.py:25-33,80: LocationInfo's input mix on its Philox stream / injected / test
    mean, and the concat / slice / add kernels around it."""
    rng = np.random.default_rng(5)
    B, seed, step, s0 = 37, 77, 4, 20
    aux = rng.random((B, 2, 2)).astype(np.float32)
    ctl = make_ctl(C, step=step, sample0=s0)
    ad = dev(aux.reshape(B, 4))
    loc = torch.zeros((B, 2), device='cuda')
    u = philox.aux_uniforms(seed, step, np.arange(s0, s0 + B)).reshape(B, 1)
    want = (aux[:, 0, :] * u + aux[:, 1, :] * (np.float32(1) - u)) * np.float32(1.5)
    C.call('tn_aux_location_mix', C.ptr(ad), C.ptr(loc), B, 1.5, 1, seed, C.ptr(ctl), None, None)
    sync()
    assert np.array_equal(loc.cpu().numpy(), want.astype(np.float32))
    u2 = rng.random(B).astype(np.float32)
    C.call('tn_aux_location_mix', C.ptr(ad), C.ptr(loc), B, 1.0, 1, 0, None, C.ptr(dev(u2)), None)
    sync()
    assert np.array_equal(loc.cpu().numpy(), aux[:, 0, :] * u2[:, None] + aux[:, 1, :] * (np.float32(1) - u2[:, None]))
    C.call('tn_aux_location_mix', C.ptr(ad), C.ptr(loc), B, 2.0, 0, 0, None, None, None)
    sync()
    assert np.array_equal(loc.cpu().numpy(), (aux[:, 0, :] + aux[:, 1, :]) / np.float32(2) * np.float32(2))
    a, c = rng.standard_normal((B, 12)).astype(np.float32), rng.standard_normal((B, 9)).astype(np.float32)
    out = torch.zeros((B, 21), device='cuda')
    C.call('tn_concat_cols', C.ptr(dev(a)), 12, C.ptr(dev(c)), 9, C.ptr(out), B, None)
    sync()
    assert np.array_equal(out.cpu().numpy(), np.concatenate([a, c], axis=1))
    back = torch.zeros((B, 9), device='cuda')
    C.call('tn_slice_cols', C.ptr(out), 21, 12, 9, C.ptr(back), B, None)
    sync()
    assert np.array_equal(back.cpu().numpy(), c)
    d1, d2 = dev(a.copy()), dev(a[::-1].copy())
    C.call('tn_add_inplace', C.ptr(d1), C.ptr(d2), a.size, None)
    sync()
    assert np.array_equal(d1.cpu().numpy(), a + a[::-1])


@pytest.mark.parametrize('planes,S,s', [(7, 12, 2), (3, 9, 3), (5, 8, 1), (2, 10, 5)])
def test_subsample_and_its_gradient(C, planes, S, s):
    """Strided ConvLayer plumbing (convpool.py:54-56): sampling lattice and its zero-filled adjoint."""
    rng = np.random.default_rng(S * s)
    O = S // s
    x = rng.standard_normal((planes, S, S)).astype(np.float32)
    out = torch.zeros((planes, O, O), device='cuda')
    C.call('tn_subsample2d', C.ptr(dev(x)), C.ptr(out), planes, S, s, O, None)
    sync()
    assert np.array_equal(out.cpu().numpy(), x[:, ::s, ::s][:, :O, :O])
    g = rng.standard_normal((planes, O, O)).astype(np.float32)
    up = torch.full((planes, S, S), 3., device='cuda')
    C.call('tn_upsample2d_zero', C.ptr(dev(g)), C.ptr(up), planes, S, s, O, None)
    sync()
    want = np.zeros((planes, S, S), np.float32)
    want[:, :O * s:s, :O * s:s] = g
    assert np.array_equal(up.cpu().numpy(), want)


@pytest.mark.parametrize('world', [1, 2, 4])
def test_peer_allreduce_two_shot_on_one_device(C, world):
    """tn_peer_allreduce with all `world` ranks living on ONE device (a rank = a buffer + a flag array +
    a stream): the kernels of the ranks run concurrently and go through the real two-phase
    handshake (reduce-scatter, all-gather).  Every rank must end with the same bits: the sum in rank
    order.  Repeated launches exercise the execution-counter tokens and the ticket reset."""
    import ctypes
    rng = np.random.default_rng(77 + world)
    off, n = 8, 4 * 3001                      # a range inside a larger buffer, ragged slices
    total = off + n + 12
    streams = [torch.cuda.Stream() for _ in range(world)]
    flags = [torch.zeros(32, dtype=torch.int32, device='cuda') for _ in range(world)]
    arr = ctypes.c_void_p * world
    for rep in range(3):
        host = [rng.standard_normal(total).astype(np.float32) for _ in range(world)]
        bufs = [dev(h) for h in host]
        want = host[0][off:off + n].copy()
        for r in range(1, world):
            want = want + host[r][off:off + n]             # float32 adds in rank order
        bp, fp = arr(*[b.data_ptr() for b in bufs]), arr(*[f.data_ptr() for f in flags])
        sync()
        for r in range(world):
            C.call('tn_peer_allreduce', bp, fp, world, r, off, n, ctypes.c_void_p(streams[r].cuda_stream))
        sync()
        for r in range(world):
            got = bufs[r].cpu().numpy()
            assert np.array_equal(got[off:off + n], want), (rep, r)
            assert np.array_equal(got[:off], host[r][:off]) and np.array_equal(got[off + n:], host[r][off + n:])
        assert all(int(f[24]) == rep + 1 and int(f[25]) == 0 for f in flags)
