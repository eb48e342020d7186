"""CPU tests of the oracle itself: independent second opinions (torch autograd, scipy, brute-force
loops) for every restated function.  The oracle is test infrastructure; these keep it honest."""
import ast
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import philox
from oracle import theanet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_prms(name):
    with open(os.path.join(ROOT, 'params', name)) as f:
        return ast.literal_eval(f.read())


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want


def test_philox_mask_statistics_and_sharding():
    m = philox.bernoulli_mask(1234, philox.PURPOSE_DROPOUT, 7, np.arange(64), 500, 0.5)
    assert abs(m.mean() - 0.5) < 0.01
    # masks are keyed by global sample index: a shard sees the same rows
    m2 = philox.bernoulli_mask(1234, philox.PURPOSE_DROPOUT, 7, np.arange(32, 64), 500, 0.5)
    assert np.array_equal(m[32:], m2)
    z = philox.elastic_noise(99, 3, 2 * 28 * 28)
    assert abs(z.mean()) < 0.1 and abs(z.std() - 1) < 0.1


@pytest.mark.parametrize("mode,f,S,C,M", [('valid', 3, 9, 2, 3), ('same', 3, 8, 3, 4),
                                          ('same', 4, 7, 2, 2), ('valid', 5, 11, 1, 2)])
def test_conv_matches_torch_flipped_kernel(mode, f, S, C, M):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, C, S, S))
    W = rng.standard_normal((M, C, f, f))
    z, cache = O.conv_forward(x, W, mode)
    xt = torch.tensor(x, requires_grad=True)
    Wt = torch.tensor(W, requires_grad=True)
    pad_lo, out_sz = O.conv_geometry(S, f, mode)
    pad_hi = out_sz + f - 1 - S - pad_lo
    zt = F.conv2d(F.pad(xt, (pad_lo, pad_hi, pad_lo, pad_hi)), Wt.flip(2, 3))
    assert np.allclose(z, zt.detach().numpy(), atol=1e-12)
    g = rng.standard_normal(z.shape)
    zt.backward(torch.tensor(g))
    dW, db, dx = O.conv_backward(g, W, cache)
    assert np.allclose(dW, Wt.grad.numpy(), atol=1e-10)
    assert np.allclose(dx, xt.grad.numpy(), atol=1e-10)
    assert np.allclose(db, g.sum(axis=(0, 2, 3)), atol=1e-10)


def test_conv_is_true_convolution_scipy():
    from scipy.signal import convolve2d
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 1, 7, 7))
    W = rng.standard_normal((1, 1, 3, 3))
    z, _ = O.conv_forward(x, W, 'valid')
    assert np.allclose(z[0, 0], convolve2d(x[0, 0], W[0, 0], mode='valid'))
    z, _ = O.conv_forward(x, W, 'same')
    assert np.allclose(z[0, 0], convolve2d(x[0, 0], W[0, 0], mode='full')[1:8, 1:8])


@pytest.mark.parametrize("S,p,ib", [(26, 2, False), (11, 2, False), (11, 2, True), (7, 3, False)])
def test_pool_bruteforce_with_ties(S, p, ib):
    rng = np.random.default_rng(2)
    x = rng.integers(-2, 3, size=(2, 3, S, S)).astype(np.float32)     # many ties
    out, cache = O.pool_forward(x, p, ib)
    o = O.pool_out_size(S, p, ib)
    assert out.shape == (2, 3, o, o)
    g = rng.standard_normal(out.shape).astype(np.float32)
    dx = O.pool_backward(g, cache)
    ref_out = np.zeros_like(out)
    ref_dx = np.zeros_like(x)
    for b in range(2):
        for c in range(3):
            for i in range(o):
                for j in range(o):
                    w = x[b, c, i * p:min((i + 1) * p, S), j * p:min((j + 1) * p, S)]
                    ref_out[b, c, i, j] = w.max()
                    hit = (w == w.max())
                    ref_dx[b, c, i * p:i * p + w.shape[0], j * p:j * p + w.shape[1]] += hit * g[b, c, i, j]
    assert np.array_equal(out, ref_out)
    assert np.array_equal(dx, ref_dx)


@pytest.mark.parametrize("name", ['relu', 'relu00', 'relu05', 'relu10', 'relu50', 'tanh',
                                  'scaled_tanh', 'sigmoid', 'softplus', 'linear'])
def test_activation_grads_match_torch(name):
    rng = np.random.default_rng(3)
    z = rng.standard_normal((5, 7))
    z[np.abs(z) < 0.05] += 0.2
    a = O.act_forward(name, z)
    g = rng.standard_normal(z.shape)
    gz = O.act_backward(name, z, a, g)
    zt = torch.tensor(z, requires_grad=True)
    if name == 'relu':
        at = torch.clamp(zt, min=0)
    elif name.startswith('relu'):
        at = torch.clamp(zt, min=0) + torch.clamp(zt, max=0) * int(name[4:]) / 100
    elif name == 'tanh':
        at = torch.tanh(zt)
    elif name == 'scaled_tanh':
        at = 1.7 * torch.tanh(2 * zt / 3)
    elif name == 'sigmoid':
        at = torch.sigmoid(zt)
    elif name == 'softplus':
        at = F.softplus(zt)
    else:
        at = zt * 1
    at.backward(torch.tensor(g))
    assert np.allclose(a, at.detach().numpy(), atol=1e-12)
    assert np.allclose(gz, zt.grad.numpy(), atol=1e-10)


def test_relu_slope_at_zero_is_both_branches():
    z = np.zeros((1, 3))
    g = np.ones((1, 3))
    assert np.allclose(O.act_backward('relu10', z, z, g), 1.1)
    assert np.allclose(O.act_backward('relu', z, z, g), 1.0)


def test_init_consumption_order_and_values():
    p = load_prms('mnist.prms')
    p['training_params']['SEED'] = 555555
    net = O.OracleNet(p['layers'], p['training_params'], img_sz=28)
    rs = np.random.RandomState(555555)
    el_seed = rs.randint(1e6)
    assert net.spec[0]['seed'] == el_seed
    w1 = (2. * rs.randint(2, size=(4, 1, 3, 3)) - 1) / np.sqrt(9)
    assert np.array_equal(net.spec[1]['params'][0], w1.astype(np.float32))
    assert np.all(net.spec[1]['params'][1] == 0)            # relu10: no +.5 bias
    w2 = (2. * rs.randint(2, size=(20, 4, 3, 3)) - 1) / np.sqrt(36)
    assert np.array_equal(net.spec[3]['params'][0], w2.astype(np.float32))
    assert np.all(net.spec[3]['params'][1] == .5)           # relu05 starts with relu0
    wh = rs.uniform(low=-1, high=1, size=(720, 500)) * np.sqrt(6 / (1220 + 1220))
    assert np.array_equal(net.spec[5]['params'][0], wh.astype(np.float32))
    assert np.all(net.spec[5]['params'][1] == .5)           # default relu01
    assert net.spec[5]['seed'] == rs.randint(1e6)
    ws = rs.uniform(low=-1, high=1, size=(500, 10)) * np.sqrt(6 / (510 + 510))
    assert np.array_equal(net.spec[6]['params'][0], ws.astype(np.float32))
    assert [L['n_out'] for L in net.spec] == [784, 2704, 676, 2420, 720, 500, 10]


def _torch_reference_step(net, x, y, masks, lr):
    """Independent torch-autograd evaluation of cost + grads for a conv/pool/dense net without
    pool ties (random data) -- checks the oracle's hand-written backward as a whole."""
    params = []
    a = torch.tensor(x, dtype=torch.float64)
    B = a.shape[0]
    for li, L in enumerate(net.spec):
        k = L['kind']
        if k == 'ConvLayer':
            W = torch.tensor(L['params'][0], dtype=torch.float64, requires_grad=True)
            b = torch.tensor(L['params'][1], dtype=torch.float64, requires_grad=True)
            params.append((li, W, b))
            f = W.shape[2]
            pad_lo, out_sz = O.conv_geometry(a.shape[2], f, L['mode'])
            pad_hi = out_sz + f - 1 - a.shape[2] - pad_lo
            z = F.conv2d(F.pad(a, (pad_lo, pad_hi, pad_lo, pad_hi)), W.flip(2, 3)) + b[None, :, None, None]
            s = int(L['actvn'][4:]) / 100
            a = torch.clamp(z, min=0) + torch.clamp(z, max=0) * s
        elif k == 'PoolLayer':
            a = F.max_pool2d(a, L['args']['pool_sz'], ceil_mode=not L['args'].get('ignore_border', False))
        elif k == 'HiddenLayer':
            W = torch.tensor(L['params'][0], dtype=torch.float64, requires_grad=True)
            b = torch.tensor(L['params'][1], dtype=torch.float64, requires_grad=True)
            params.append((li, W, b))
            z = a.reshape(B, -1) @ W + b
            s = int(L['actvn'][4:]) / 100
            a = torch.clamp(z, min=0) + torch.clamp(z, max=0) * s
            if L['pdrop']:
                a = a * torch.tensor(masks[li], dtype=torch.float64)
        elif k == 'SoftmaxLayer':
            W = torch.tensor(L['params'][0], dtype=torch.float64, requires_grad=True)
            b = torch.tensor(L['params'][1], dtype=torch.float64, requires_grad=True)
            params.append((li, W, b))
            a = F.log_softmax(a.reshape(B, -1) @ W + b, dim=1)
    cost = -a[torch.arange(B), torch.tensor(y)].mean()
    for li, W, b in params:
        reg = net.spec[li]['reg']
        cost = cost + reg['L1'] * (W.abs().sum() + b.abs().sum()) + reg['L2'] * ((W ** 2).sum() + (b ** 2).sum())
    cost.backward()
    return cost.item(), {li: (W.grad.numpy(), b.grad.numpy()) for li, W, b in params}


def test_whole_step_gradients_match_torch_autograd():
    layers = [('InputLayer', {'img_sz': 12, 'num_maps': 2}),
              ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu10',
                             'reg': {'L1': 1e-3, 'L2': 1e-2}}),
              ('PoolLayer', {'pool_sz': 2}),
              ('ConvLayer', {'num_maps': 5, 'filter_sz': 3, 'stride': 1, 'mode': 'same', 'actvn': 'relu05'}),
              ('PoolLayer', {'pool_sz': 2}),
              ('HiddenLayer', {'n_out': 17, 'pdrop': .5, 'reg': {'L2': .01}}),
              ('SoftmaxLayer', {'n_out': 6})]
    tp = {'SEED': 7, 'BATCH_SZ': 8, 'INIT_LEARNING_RATE': .1, 'EPOCHS_TO_HALF_RATE': 1}
    net = O.OracleNet(layers, tp, dtype=np.float64)
    rng = np.random.default_rng(5)
    for L in net.spec:       # perturb so L1 sign(theta) and biases are generic
        L['params'] = [p + 0.01 * rng.standard_normal(p.shape) for p in L['params']]
    x = rng.standard_normal((8, 2, 12, 12))
    y = rng.integers(0, 6, 8)
    mask = (rng.random((8, 17)) < .5).astype(np.float64)
    cost, _ = net.train_step(x, y, rand={(5, 'mask'): mask}, apply_update=False)
    tcost, tg = _torch_reference_step(net, x, y, {5: mask}, 0.1)
    assert abs(cost - tcost) < 1e-10
    for li, (gW, gb) in tg.items():
        reg = net.spec[li]['reg']
        W, b = net.spec[li]['params']
        oW = net.last_grads[li][0] + reg['L1'] * np.sign(W) + 2 * reg['L2'] * W
        ob = net.last_grads[li][1] + reg['L1'] * np.sign(b) + 2 * reg['L2'] * b
        assert np.allclose(oW, gW, atol=1e-10), li
        assert np.allclose(ob, gb, atol=1e-10), li


def test_lagged_momentum_first_step_moves_nothing():
    p = load_prms('3flat.prms')
    p['training_params']['SEED'] = 555555
    net = O.OracleNet(p['layers'], p['training_params'], img_sz=28)
    rng = np.random.default_rng(0)
    x = rng.random((20, 1, 28, 28), dtype=np.float32)
    y = rng.integers(0, 457, 20)
    before = [q.copy() for q in net.spec[1]['params']]
    net.train_step(x, y, step=0)
    assert all(np.array_equal(a, b) for a, b in zip(before, net.spec[1]['params']))
    assert np.abs(net.spec[1]['vel'][0]).max() > 0
    net.train_step(x, y, step=1)
    assert not np.array_equal(before[0], net.spec[1]['params'][0])


def test_maxnorm_variants():
    rng = np.random.default_rng(0)
    reg = dict(O.DEFAULT_REG, maxnorm=0.5)
    th = rng.standard_normal((6, 4)).astype(np.float32)
    th[:, 2] = 0            # zero-norm column -> scale (1e-7)/(1e-7) = 1
    new, _ = O.sgd_update(th, np.zeros_like(th), np.zeros_like(th), reg, 0.1)
    n = np.sqrt((new ** 2).sum(0))
    assert np.all(n <= 0.5 + 1e-5) and np.all(new[:, 2] == 0)
    th4 = rng.standard_normal((3, 2, 3, 3)).astype(np.float32)
    new4, _ = O.sgd_update(th4, np.zeros_like(th4), np.zeros_like(th4), reg, 0.1)
    assert np.all(np.sqrt((new4 ** 2).sum((1, 2, 3))) <= 0.5 + 1e-5)
    b = np.array([-2, .1, 3], np.float32)
    newb, _ = O.sgd_update(b, np.zeros_like(b), np.zeros_like(b), reg, 0.1)
    assert np.array_equal(newb, np.array([-.5, .1, .5], np.float32))


def test_elastic_smoothing_matches_scipy_full_conv():
    from scipy.signal import convolve2d
    rng = np.random.default_rng(4)
    h, sigma = 12, 3
    noise = rng.standard_normal((2, h, h)).astype(np.float32)
    prm = {'magnitude': 7, 'sigma': sigma}
    ty, tx, disp = O.elastic_target(h, prm, noise, np.full(8, .5, np.float32))
    filt = O.gaussian_filter(sigma).astype(np.float64)
    for c in range(2):
        full = convolve2d((np.float32(7) * noise[c]).astype(np.float64), filt, mode='full')
        want = full[sigma:h + sigma, sigma:h + sigma].astype(np.float32).astype(np.float64)
        assert np.allclose(disp[c], want, atol=1e-12)


def test_elastic_nearest_bilinear_and_flip():
    rng = np.random.default_rng(6)
    h = 10
    x = rng.random((3, 2, h, h), dtype=np.float32)
    prm = {'translation': 2, 'zoom': 1.2, 'angle': 10, 'invert_image': True, 'nearest': True}
    u = rng.random(8).astype(np.float32)
    ty, tx, _ = O.elastic_target(h, prm, None, u)
    assert ty.min() >= 0 and ty.max() <= h - 1 - .001
    out = O.elastic_apply(x, prm, ty, tx)
    v, hz = np.rint(np.floor(ty + .5)).astype(int), np.floor(tx + .5).astype(int)
    assert np.array_equal(out, (1 - x)[:, :, v, hz])
    prm['nearest'] = False
    out = O.elastic_apply(x, prm, ty, tx)
    assert out.shape == x.shape and out.min() >= 0 and out.max() <= 1
    m = (rng.random(x.shape) < .3).astype(np.float32)
    outf = O.elastic_apply(x, prm, ty, tx, m)
    assert np.array_equal(outf[m == 1], (1 - out)[m == 1]) and np.array_equal(outf[m == 0], out[m == 0])
    # identity shortcut: only the inversion survives
    assert O.elastic_is_identity({'invert_image': True, 'nearest': True})


def test_test_twin_scales_dropout():
    layers = [('InputLayer', {'img_sz': 4}), ('HiddenLayer', {'n_out': 5, 'pdrop': .25}),
              ('SoftmaxLayer', {'n_out': 3})]
    tp = {'SEED': 1, 'BATCH_SZ': 2, 'INIT_LEARNING_RATE': .1, 'EPOCHS_TO_HALF_RATE': 1}
    net = O.OracleNet(layers, tp, dtype=np.float64)
    x = np.random.default_rng(0).random((2, 1, 4, 4))
    err, pm, logprob, preds = net.test_step(x, [0, 1])
    W, b = net.spec[1]['params']
    h = O.act_forward('relu01', x.reshape(2, -1) @ W + b) * .75
    W2, b2 = net.spec[2]['params']
    assert np.allclose(logprob, O.log_softmax(h @ W2 + b2))
    assert np.array_equal(preds, logprob.argmax(1))


def test_bf16_rounding_matches_torch_and_eligibility():
    """The oracle's bf16 emulation (NOT reference behaviour: theanet_b200's optional mixed-precision
    conv stack, config C4) rounds exactly like torch / the GPU (round to nearest even)."""
    import torch
    rng = np.random.default_rng(3)
    x = np.concatenate([(rng.standard_normal(100000) * 10.0 ** rng.integers(-6, 6, 100000)).astype(np.float32),
                        np.array([1.00390625, 1.01171875, -1.00390625, 0., 65504.], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(O.bf16_round(x), want)
    assert O.conv_tc_eligible(64, 16, 128, 3, 'same', 'relu05', 16, 2)
    assert O.conv_tc_eligible(3, 32, 64, 3, 'same', 'relu05', 32, 2, first=True)       # im2col
    assert not O.conv_tc_eligible(3, 32, 64, 3, 'same', 'relu05', 32, 2, first=False)
    assert not O.conv_tc_eligible(64, 16, 128, 3, 'valid', 'relu05', 14, 2)
    assert not O.conv_tc_eligible(64, 16, 128, 3, 'same', 'tanh', 16, 2)
    assert not O.conv_tc_eligible(64, 28, 128, 3, 'same', 'relu05', 28, 2)              # 28 does not tile


def test_bf16_conv_stack_mode_changes_only_eligible_layers():
    prms = load_prms('mnist.prms')
    prms['training_params'].update(SEED=5, BATCH_SZ=4, CONV_DTYPE='bfloat16')
    prms['layers'][0][1]['img_sz'] = 28
    on = O.OracleNet(prms['layers'], prms['training_params'])
    assert not any(L['tc'] for L in on.spec)        # 'valid' 4/20-map convs stay float32
    p2 = ast.literal_eval(open(os.path.join(ROOT, 'params', 'cifar3conv.prms')).read())
    p2['training_params'].update(SEED=5, BATCH_SZ=2, CONV_DTYPE='bfloat16')
    p2['layers'][0][1]['img_sz'] = 32
    on2 = O.OracleNet(p2['layers'], p2['training_params'])
    assert [L['tc'] for L in on2.spec if L['kind'] == 'ConvLayer'] == [True, True, True]
    rng = np.random.default_rng(0)
    x = rng.random((2, 3, 32, 32), dtype=np.float32)
    cost, lp = on2.train_step(x, np.array([1, 7]), step=0)
    assert np.isfinite(cost) and lp.shape == (2, 10)
    a = on2._forward(x, False, 0, 0)[1][1]['a']     # conv 1 activations are bf16 values
    assert np.array_equal(O.bf16_round(a), a)


@pytest.mark.parametrize('kind,loss', [('SoftmaxLayer', 'nll'), ('SoftmaxLayer', 'nllsq'),
                                       ('SoftmaxLayer', 'nll35'), ('ExpLossLayer', 'exp'),
                                       ('HingeLayer', 'hinge')])
def test_output_losses_match_torch_autograd(kind, loss):
    """output_views / output_loss (outlayers.py:38-64,105-147) against torch autograd on the
    formulas as the reference writes them (off the kinks: the tie rule A5 is pinned elsewhere)."""
    import torch
    rng = np.random.default_rng(11)
    B, n, Bg = 9, 6, 18
    z = rng.standard_normal((B, n))
    y = rng.integers(0, n, B)
    feat, lp, probs = O.output_views(kind, z)
    cost, g = O.output_loss(kind, loss, z, lp, y, Bg)
    zt = torch.tensor(z, requires_grad=True)
    rows = torch.arange(B)
    yt = torch.tensor(y)
    if kind == 'SoftmaxLayer':
        lpt = torch.log(torch.softmax(zt, dim=1))
        pick = lpt[rows, yt]
        if loss == 'nll':
            per = -pick
        elif loss == 'nllsq':
            per = pick ** 2
        else:
            per = torch.clamp(np.log(.35) - pick, min=0)
        want_feat = lpt
    elif kind == 'ExpLossLayer':
        o = zt - zt.mean(dim=1, keepdim=True)
        per = torch.exp(-o[rows, yt])
        want_feat = o
    else:
        per = torch.clamp(zt + 1 - zt[rows, yt][:, None], min=0).sum(dim=1) / n
        want_feat = zt
    total = per.sum() / Bg
    total.backward()
    assert abs(cost - total.item()) < 1e-12
    assert np.allclose(g, zt.grad.numpy(), atol=1e-12)
    assert np.allclose(feat, want_feat.detach().numpy(), atol=1e-12)


def test_meanlayer_colorlayer_and_aux_restatements():
    import torch
    rng = np.random.default_rng(12)
    # ColorLayer (color.py:36-44), written out independently in float64
    x = rng.random((3, 2, 4, 4)).astype(np.float32) * 2
    u = rng.uniform(-1, 1, (3, 2, 3)).astype(np.float32)
    prm = {'balance': 1.4, 'gamma': 1.7, 'maxval': 2}
    out = O.color_jitter(x, prm, u)
    e = [np.exp(np.log(a) * u[:, :, k].astype(np.float64))[:, :, None, None]
         for k, a in enumerate((1.4, 1.7, 1.7))]
    w = np.clip(x.astype(np.float64) / 2 * e[0], 0, 1) ** e[1]
    want = (1 - (1 - w) ** e[2]) * 2
    assert np.max(np.abs(out - want)) < 2e-6
    # a net with a MeanLayer: whole-step gradients against torch autograd
    layers = [('InputLayer', {'img_sz': 6, 'num_maps': 2}),
              ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'mode': 'same', 'actvn': 'tanh'}),
              ('MeanLayer', {}),
              ('SoftmaxLayer', {'n_out': 4})]
    tp = {'SEED': 3, 'BATCH_SZ': 5, 'INIT_LEARNING_RATE': .1, 'EPOCHS_TO_HALF_RATE': 1}
    on = O.OracleNet(layers, tp, dtype=np.float64)
    xb = rng.random((5, 2, 6, 6))
    yb = rng.integers(0, 4, 5)
    cost, _ = on.train_step(xb, yb, apply_update=False)
    W, b = (torch.tensor(t, requires_grad=True) for t in on.spec[1]['params'])
    Ws, bs = (torch.tensor(t, requires_grad=True) for t in on.spec[3]['params'])
    a = torch.tanh(torch.nn.functional.conv2d(torch.tensor(xb), torch.flip(W, (2, 3)), padding=1)
                   + b[None, :, None, None])
    z = a.mean(dim=(2, 3)) @ Ws + bs
    loss = -torch.log_softmax(z, dim=1)[torch.arange(5), torch.tensor(yb)].mean()
    loss.backward()
    assert abs(cost - loss.item()) < 1e-12
    for got, want_t in zip(on.last_grads[1] + on.last_grads[3], (W, b, Ws, bs)):
        assert np.allclose(got, want_t.grad.numpy(), atol=1e-12)


def test_kink_localisation_hook_only_touches_elements_at_the_kink():
    """OracleNet.kink_source (tests/test_gpu_net.py::run_pair): with the oracle's OWN hidden output as
    the source nothing is re-signed; flipping the sign of one tiny pre-activation in the source is
    recorded as exactly one flip whose |z| is what it was, and moves only that unit's gradient."""
    import copy
    layers = [('InputLayer', {'img_sz': 6, 'num_maps': 1}),
              ('HiddenLayer', {'n_out': 16, 'pdrop': 0, 'actvn': 'relu10'}),
              ('SoftmaxLayer', {'n_out': 5})]
    tp = {'BATCH_SZ': 8, 'NUM_EPOCHS': 1, 'EPOCHS_TO_TEST': 1, 'TEST_SAMP_SZ': 8, 'INIT_LEARNING_RATE': .1,
          'EPOCHS_TO_HALF_RATE': 1, 'SEED': 3}
    rng = np.random.default_rng(0)
    x = rng.standard_normal((8, 1, 6, 6)).astype(np.float32)
    y = rng.integers(0, 5, 8)
    a = O.OracleNet(copy.deepcopy(layers), dict(tp))
    a.train_step(x, y, step=0, apply_update=False)
    base = [g.copy() for g in a.last_grads[1]]
    out = a.last_caches[1]['out'].copy()
    z = a.last_caches[1]['z']
    b = O.OracleNet(copy.deepcopy(layers), dict(tp))
    b.kink_source = {1: out}
    b.train_step(x, y, step=0, apply_update=False)
    assert b.kink_flips == [(1, 0, 0.0, float(np.abs(z).max()))]
    assert all(np.array_equal(u, v) for u, v in zip(base, b.last_grads[1]))
    # the other implementation landed on the other side of zero for the smallest |z|
    i, j = np.unravel_index(np.argmin(np.abs(z)), z.shape)
    other = out.copy()
    other[i, j] = -np.sign(z[i, j]) * 1e-9
    c = O.OracleNet(copy.deepcopy(layers), dict(tp))
    c.kink_source = {1: other}
    c.train_step(x, y, step=0, apply_update=False)
    (li, n, zflip, zmax), = c.kink_flips
    assert (li, n) == (1, 1) and np.isclose(zflip, abs(z[i, j])) and zmax == float(np.abs(z).max())
    dW = c.last_grads[1][0] - base[0]
    assert np.any(dW[:, j] != 0) and not np.any(np.delete(dW, j, axis=1))
