"""CPU restatement (numpy) of theanet's CNN-training hot path -- TEST INFRASTRUCTURE ONLY.

PARITY STATUS: the reference (rakeshvar/theanet) ships no golden vectors or asserting tests
(tests/test_elastic.py is a visual script) and its engine, Theano, is absent from this
environment and cannot be installed (Python 3.12 / NumPy 2.3; no network).  This restatement
follows the reference *source* line by line (citations below) and is pinned against the
reference's own, unmodified Python executed over a stand-in for the Theano API
(oracle/theano_shim; generator tests/golden/make_golden_ref.py; check tests/test_golden_ref.py):
initial weights bit-identical, costs / log-probabilities / updated weights / elastic fields to
~1e-5 over three networks.  That pins everything the reference itself writes; upstream Theano op
semantics (assumptions A1-A9) are encoded identically in shim and oracle and remain
unverified against a real Theano -- for those items parity is "unpinned" (DESIGN.md section 2).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may
import this module; the product path (theanet_b200/) never does.

Conventions: activations NCHW, conv weights OIHW, dense weights (n_in, n_out) -- the layouts
the reference's .pkl exposes (theanet/neuralnet.py:298-301).  ``dtype`` selects float32 (the
reference's floatX) or float64 ("truth") arithmetic.
"""
import math
import numpy as np

from . import philox

# ----------------------------------------------------------------------------------------------
# Activations -- theanet/layer/layer.py:27-54
# ----------------------------------------------------------------------------------------------
ACT_NAMES = (['sigmoid', 'softplus', 'softmax', 'linear', 'scaled_tanh', 'relu', 'tanh'] +
             ['relu{:02d}'.format(i) for i in range(100)])


def act_forward(name, z):
    """z -> act(z) in z's dtype.  reluNN follows layer.py:36 exactly: max(0,x) + (min(0,x)*NN)/100."""
    dt = z.dtype.type
    if name in ('linear',):
        return z
    if name == 'relu':
        return np.maximum(dt(0), z)
    if name.startswith('relu') and len(name) == 6:
        nn = dt(int(name[4:]))
        return np.maximum(dt(0), z) + (np.minimum(dt(0), z) * nn) / dt(100)
    if name == 'tanh':
        return np.tanh(z)
    if name == 'scaled_tanh':
        return dt(1.7) * np.tanh(dt(2) * z / dt(3))
    if name == 'sigmoid':
        return dt(1) / (dt(1) + np.exp(-z))
    if name == 'softplus':
        return np.logaddexp(dt(0), z)
    raise NotImplementedError("Unknown Activation Specified: " + name)


def act_backward(name, z, a, g):
    """dL/dz from dL/da = g.  At z == 0 the maximum/minimum gradients both fire (assumption A5:
    Theano's Maximum.grad uses eq(out, arg)), so reluNN has slope 1 + NN/100 there, relu has 1."""
    dt = z.dtype.type
    if name == 'linear':
        return g
    if name == 'relu':
        return g * (z >= 0).astype(z.dtype)
    if name.startswith('relu') and len(name) == 6:
        s = dt(int(name[4:])) / dt(100)
        d = np.where(z > 0, dt(1), np.where(z < 0, s, dt(1) + s)).astype(z.dtype)
        return g * d
    if name == 'tanh':
        return g * (dt(1) - a * a)
    if name == 'scaled_tanh':
        t = a / dt(1.7)
        return g * (dt(1.7) * dt(2) / dt(3)) * (dt(1) - t * t)
    if name == 'sigmoid':
        return g * a * (dt(1) - a)
    if name == 'softplus':
        return g * (dt(1) - np.exp(-a))
    raise NotImplementedError(name)


# ----------------------------------------------------------------------------------------------
# Weight init -- theanet/layer/weights.py:25-81
# ----------------------------------------------------------------------------------------------
def init_wb(rand_gen, size_w, size_b, fan_in, fan_out, actvn):
    if len(size_w) == 4:                                   # weights.py:51-54
        w = 2. * rand_gen.randint(2, size=size_w) - 1
        w /= np.sqrt(fan_in)
    else:                                                  # weights.py:56-57
        w = rand_gen.uniform(low=-1, high=1, size=size_w)
        w *= np.sqrt(6 / (fan_in + fan_out))
    w = np.asarray(w, dtype=np.float32)                    # weights.py:59
    b = np.zeros(size_b, dtype=np.float32)
    if actvn == 'sigmoid':                                 # weights.py:62-63
        w *= 4
    if actvn in ('softplus', 'relu') or actvn.startswith('relu0'):   # weights.py:64-65
        b += .5
    return w, b


# ----------------------------------------------------------------------------------------------
# Convolution (true convolution, kernel flipped) -- theanet/layer/convpool.py:42-72, A1
# ----------------------------------------------------------------------------------------------
def conv_geometry(in_sz, f, mode):
    """(pad_lo, out_sz) for stride 1.  'same' = 'full' cropped by (f-1)//2 (convpool.py:57-61)."""
    if mode == 'valid':
        return 0, in_sz - f + 1
    if mode == 'same':
        shift = (f - 1) // 2
        return f - 1 - shift, in_sz
    raise NotImplementedError("mode='full' has a wrong out_sz in the reference (convpool.py:64)")


def _im2col(xp, f, out_sz):
    """xp (B,C,Sp,Sp) padded -> cols (C*f*f, B*out*out), a copy."""
    B, C = xp.shape[:2]
    s = xp.strides
    v = np.lib.stride_tricks.as_strided(
        xp, shape=(C, f, f, B, out_sz, out_sz),
        strides=(s[1], s[2], s[3], s[0], s[2], s[3]), writeable=False)
    return np.ascontiguousarray(v).reshape(C * f * f, B * out_sz * out_sz)


# Accumulation type of the forward convolution: None = the working dtype (float32 BLAS GEMM, what
# Theano's CorrMM does); np.float64 = accumulate in double and round once, which makes the result
# independent of the summation order.  It matters only for max-pool ties (A3) between sums that
# are equal mathematically but not operand for operand (binary +-c conv weights over duplicated or
# flat non-zero pixels): there each float32 implementation -- BLAS, cuDNN-style kernels, ours --
# breaks or keeps the tie according to its own summation order, and no order is canonical.  Tests
# therefore use data whose ties are exact in any order (zero background); see DESIGN.md section 2.
CONV_ACCUM = None


def conv_forward(x, W, mode='valid'):
    """out[b,m,i,j] = sum_{c,u,v} xpad[b,c,i+u,j+v] * W[m,c,f-1-u,f-1-v]; returns (z, cache)."""
    B, C, S, _ = x.shape
    M, _, f, _ = W.shape
    pad_lo, out_sz = conv_geometry(S, f, mode)
    pad_hi = out_sz + f - 1 - S - pad_lo
    xp = np.pad(x, ((0, 0), (0, 0), (pad_lo, pad_hi), (pad_lo, pad_hi))) if (pad_lo or pad_hi) else x
    cols = _im2col(xp, f, out_sz)
    Wf = W[:, :, ::-1, ::-1].reshape(M, C * f * f)
    if CONV_ACCUM is not None and np.dtype(CONV_ACCUM) != x.dtype:
        z = (Wf.astype(CONV_ACCUM) @ cols.astype(CONV_ACCUM)).astype(x.dtype)
    else:
        z = Wf @ cols
    z = z.reshape(M, B, out_sz, out_sz).transpose(1, 0, 2, 3)
    return np.ascontiguousarray(z), (cols, x.shape, pad_lo, pad_hi, out_sz)


def conv_backward(dz, W, cache, need_dx=True):
    cols, xshape, pad_lo, pad_hi, out_sz = cache
    B, C, S, _ = xshape
    M, _, f, _ = W.shape
    dz2 = dz.transpose(1, 0, 2, 3).reshape(M, B * out_sz * out_sz)
    dWf = (dz2 @ cols.T).reshape(M, C, f, f)
    dW = np.ascontiguousarray(dWf[:, :, ::-1, ::-1])
    db = dz2.sum(axis=1)
    dx = None
    if need_dx:
        Wf = W[:, :, ::-1, ::-1].reshape(M, C * f * f)
        dcols = (Wf.T @ dz2).reshape(C, f, f, B, out_sz, out_sz)
        Sp = S + pad_lo + pad_hi
        dxp = np.zeros((B, C, Sp, Sp), dz.dtype)
        for u in range(f):
            for v in range(f):
                dxp[:, :, u:u + out_sz, v:v + out_sz] += dcols[:, u, v].transpose(1, 0, 2, 3)
        dx = np.ascontiguousarray(dxp[:, :, pad_lo:pad_lo + S, pad_lo:pad_lo + S])
    return dW, db.astype(dz.dtype), dx


# ----------------------------------------------------------------------------------------------
# Max pool -- theanet/layer/convpool.py:97-112, A2 (ceil mode) and A3 (ties all get the gradient)
# ----------------------------------------------------------------------------------------------
def pool_out_size(in_sz, p, ignore_border=False):
    return in_sz // p if ignore_border else int(math.ceil(in_sz / p))


def pool_forward(x, p, ignore_border=False):
    B, C, S, _ = x.shape
    o = pool_out_size(S, p, ignore_border)
    if ignore_border:
        xp = x[:, :, :o * p, :o * p]
    else:
        padn = o * p - S
        xp = np.pad(x, ((0, 0), (0, 0), (0, padn), (0, padn)), constant_values=-np.inf) if padn else x
    out = None                      # max over the p*p window offsets (exact, order-independent)
    for di in range(p):
        for dj in range(p):
            v = xp[:, :, di::p, dj::p]
            out = v.copy() if out is None else np.maximum(out, v, out=out)
    return out, (xp, out, S, p, o)


def pool_backward(dout, cache):
    xp, out, S, p, o = cache
    B, C = xp.shape[:2]
    win = xp.reshape(B, C, o, p, o, p)
    hit = (win == out[:, :, :, None, :, None])
    dxp = (hit * dout[:, :, :, None, :, None]).astype(dout.dtype).reshape(B, C, o * p, o * p)
    dx = np.zeros((B, C, S, S), dout.dtype)
    n = min(S, o * p)
    dx[:, :, :n, :n] = dxp[:, :, :n, :n]
    return dx


# ----------------------------------------------------------------------------------------------
# Softmax / NLL -- theanet/layer/outlayers.py:50-51,69-80,87-93, A4
# ----------------------------------------------------------------------------------------------
def log_softmax(z):
    m = z.max(axis=1, keepdims=True)
    e = z - m
    return e - np.log(np.exp(e).sum(axis=1, keepdims=True))


# ----------------------------------------------------------------------------------------------
# Elastic distortion -- theanet/layer/inlayers.py:29-163
# ----------------------------------------------------------------------------------------------
def gaussian_filter(sigma):
    """inlayers.py:87-91: float32 table, truncated at +-sigma, divided by 2*pi*var (not renormalised)."""
    var = sigma ** 2
    filt = np.array([[np.exp(-.5 * (i * i + j * j) / var)
                      for i in range(-sigma, sigma + 1)]
                     for j in range(-sigma, sigma + 1)], dtype=np.float32)
    filt /= np.float32(2 * np.pi * var)
    return filt


def elastic_target(h, prm, noise, u):
    """The per-minibatch sampling grid (inlayers.py:77-122).

    noise: (2,h,h) float32 standard normals (srs.normal, :94); u: 8 float32 uniforms in (0,1)
    [trans_y, trans_x, origin_y, origin_x, zoom_y, zoom_x, angle, spare] mapped to the reference's
    ranges here.  Dtype flow (assumption A9): random draws and the scalars derived from them
    (translation, zoom factor, cos/sin) are float32 values (Theano floatX); the grid itself is
    float64 because np.indices is int64 (:77) and int64 (+) float32 upcasts to float64.
    Transcendentals are evaluated in float64 and rounded to float32 (= a correctly rounded
    float32 libm).  Returns (transy, transx, displacement) in float64.

    ``u`` may instead be a dict of draws ALREADY mapped to the reference's ranges -- 'translation'
    and 'zoom' (2 values each, U(-1,1)), 'origin' (2 values, U(.25,.75)), 'angle' (1 value,
    U(-1,1)) -- which is how tests/golden/make_golden_ref.py hands over the float32 values the
    reference's own srs.uniform calls produced (re-deriving a (0,1) uniform would lose low bits).
    """
    w = h
    f32, f64 = np.float32, np.float64
    if isinstance(u, dict):
        d = {k: np.asarray(v, f32).reshape(-1) for k, v in u.items()}
        z2 = np.zeros(2, f32)
        m_tr, m_or = d.get('translation', z2), d.get('origin', z2)
        m_zo, m_an = d.get('zoom', z2), d.get('angle', z2)[0]
    else:
        u = np.asarray(u, f32)
        m_tr, m_or = f32(2) * u[0:2] - f32(1), f32(.25) + f32(.5) * u[2:4]
        m_zo, m_an = f32(2) * u[4:6] - f32(1), f32(2) * u[6] - f32(1)
    target = np.indices((h, w)).astype(f64)
    if prm.get('translation', 0):                                           # :80-82
        t = f32(prm['translation']) * m_tr
        target = target + t.astype(f64).reshape(2, 1, 1)
    if prm.get('magnitude', 0):                                             # :85-97
        sigma = int(prm.get('sigma', 1))
        filt = gaussian_filter(sigma).astype(f64)
        el = (f32(prm['magnitude']) * np.asarray(noise, f32)).astype(f64)
        pad = np.pad(el, ((0, 0), (sigma, sigma), (sigma, sigma)))
        k = 2 * sigma + 1
        s = pad.strides
        win = np.lib.stride_tricks.as_strided(pad, shape=(2, h, w, k, k),
                                              strides=(s[0], s[1], s[2], s[1], s[2]))
        # 'full' convolution cropped [sigma:h+sigma] == correlation with the (symmetric) table
        elast = np.einsum('chwij,ij->chw', win, filt[::-1, ::-1])
        target = target + elast.astype(f32).astype(f64)
    zoom, angle = prm.get('zoom', 1), prm.get('angle', 0)
    if zoom - 1 or angle:                                                   # :100-118
        origin = m_or.astype(f64) * np.array((h, w), f64)
        origin = origin.reshape(2, 1, 1)
        target = target - origin
        if zoom - 1:
            e = f64(f32(np.log(zoom))) * m_zo.astype(f64)
            zoomer = np.exp(e).astype(f32).astype(f64)
            target = target * zoomer.reshape(2, 1, 1)
        if angle:
            theta = f32(f32(angle * np.pi / 180) * m_an)
            c = f64(f32(np.cos(f64(theta))))
            s_ = f64(f32(np.sin(f64(theta))))
            # tensordot(rotate, target, axes=(0,0)) = R^T . target, R = [[c,-s],[s,c]]   (:113-115)
            t0 = c * target[0] + s_ * target[1]
            t1 = -s_ * target[0] + c * target[1]
            target = np.stack([t0, t1])
        target = target + origin
    transy = np.clip(target[0], 0, h - 1 - .001)                            # :121-122
    transx = np.clip(target[1], 0, w - 1 - .001)
    disp = target - np.indices((h, w))
    return transy, transx, disp


def _iround(a):
    """tt.iround = round half away from zero -> int64 (A6)."""
    return np.where(a >= 0, np.floor(a + 0.5), np.ceil(a - 0.5)).astype(np.int64)


def elastic_apply(x, prm, transy, transx, flip_mask=None):
    """Invert, gather (nearest / bilinear), pixel-flip noise (inlayers.py:63-64,124-142)."""
    dt = x.dtype.type
    if prm.get('invert_image', False):
        x = dt(1) - x
    if transy is not None:
        if prm.get('nearest', False):
            vert, horz = _iround(transy), _iround(transx)
            out = x[:, :, vert, horz]
        else:
            topp = transy.astype(np.int32)
            left = transx.astype(np.int32)
            fy = (transy - topp).astype(x.dtype)
            fx = (transx - left).astype(x.dtype)
            one = dt(1)
            out = (x[:, :, topp, left] * (one - fy) * (one - fx) +
                   x[:, :, topp, left + 1] * (one - fy) * fx +
                   x[:, :, topp + 1, left] * fy * (one - fx) +
                   x[:, :, topp + 1, left + 1] * fy * fx)
    else:
        out = x
    if flip_mask is not None:
        m = flip_mask.astype(x.dtype)
        out = (dt(1) - out) * m + out * (dt(1) - m)
    return out


def elastic_is_identity(prm):
    """inlayers.py:67-70."""
    return (not (prm.get('magnitude', 0) or prm.get('translation', 0) or prm.get('pflip', 0)
                 or prm.get('angle', 0))) and prm.get('zoom', 1) == 1


# ----------------------------------------------------------------------------------------------
# Optimiser -- theanet/layer/layer.py:70-117, A7 (simultaneous updates => lagged momentum)
# ----------------------------------------------------------------------------------------------
DEFAULT_REG = {"L1": 0, "L2": 0, "momentum": .95, "rate": 1, "maxnorm": 0}


def sgd_update(theta, vel, grad, reg, lr):
    """Returns (theta_new, vel_new).  grad must already include the L1/L2 terms."""
    dt = theta.dtype.type
    m = dt(reg['momentum'])
    vel_new = m * vel + (dt(1.) - m) * grad                                  # layer.py:82-84
    theta_new = theta - dt(reg['rate']) * dt(lr) * vel                         # layer.py:86 (OLD vel)
    mx = reg['maxnorm']
    if mx:
        mx = dt(mx)
        eps = dt(1e-7)
        if theta.ndim == 1:                                                  # layer.py:90-91
            theta_new = np.clip(theta_new, -mx, mx)
        elif theta.ndim == 2:                                                # layer.py:93-97
            n = np.sqrt(np.sum(theta_new * theta_new, axis=0))
            theta_new = theta_new * ((eps + np.clip(n, 0, mx)) / (eps + n))
        elif theta.ndim == 4:                                                # layer.py:99-103
            n = np.sqrt(np.sum(theta_new * theta_new, axis=(1, 2, 3)))
            theta_new = theta_new * ((eps + np.clip(n, 0, mx)) / (eps + n))[:, None, None, None]
    return theta_new.astype(theta.dtype), vel_new.astype(theta.dtype)


# ----------------------------------------------------------------------------------------------
# Randomness providers
# ----------------------------------------------------------------------------------------------
class PhiloxRandom:
    """The product's own streams (oracle/philox.py == csrc/philox.cuh)."""

    def elastic(self, seed, step, h):
        return philox.elastic_noise(seed, step, 2 * h * h).reshape(2, h, h), \
            philox.elastic_scalars(seed, step)

    def flip_mask(self, seed, step, samples, n, p):
        return philox.bernoulli_mask(seed, philox.PURPOSE_FLIP, step, samples, n, p)

    def keep_mask(self, seed, step, samples, n, p_keep):
        return philox.bernoulli_mask(seed, philox.PURPOSE_DROPOUT, step, samples, n, p_keep)


class InjectedRandom:
    """Randomness supplied by the test: dict keyed by (layer_index, kind) -> callable(step) or array."""

    def __init__(self, table):
        self.table = table

    def _get(self, key, step):
        v = self.table[key]
        return v(step) if callable(v) else v


# ----------------------------------------------------------------------------------------------
# ColorLayer -- theanet/layer/color.py:9-52
# ----------------------------------------------------------------------------------------------
def color_jitter(x, prm, u):
    """u: (B, maps, 3) float32 draws in (-1,1) -- balance, gamma, gamma (pos_rand is called three
    times, color.py:32-42).  np.log(a) enters the graph as a floatX constant; everything float32."""
    f32 = np.float32
    maxval = f32(prm.get('maxval', 1))
    lb, lg = f32(np.log(prm.get('balance', 1))), f32(np.log(prm.get('gamma', 1)))
    u = np.asarray(u, f32)
    e1 = np.exp(lb * u[:, :, 0])[:, :, None, None]
    e2 = np.exp(lg * u[:, :, 1])[:, :, None, None]
    e3 = np.exp(lg * u[:, :, 2])[:, :, None, None]
    out = x.astype(f32) / maxval
    out = out * e1
    out = np.clip(out, f32(0), f32(1))
    out = out ** e2
    out = f32(1) - (f32(1) - out) ** e3
    return (out * maxval).astype(x.dtype)


# ----------------------------------------------------------------------------------------------
# Output layers and losses -- theanet/layer/outlayers.py:12-147
# ----------------------------------------------------------------------------------------------
OUT_LAYERS = ('SoftmaxLayer', 'ExpLossLayer', 'HingeLayer')


def output_views(kind, z):
    """(features, logprob, probs) of an output layer from its scores z = x.w + b."""
    if kind == 'SoftmaxLayer':                       # :83-102 (A4: stable log-softmax)
        lp = log_softmax(z)
        return lp, lp, np.exp(lp)
    if kind == 'ExpLossLayer':                       # :105-126: rows centred, then softmax
        o = z - z.mean(axis=1, keepdims=True, dtype=z.dtype)
        lp = log_softmax(o)
        return o, lp, np.exp(lp)
    if kind == 'HingeLayer':                         # :129-147: logprob = probs = features = z
        return z, z, z
    raise NotImplementedError(kind)


def loss_threshold(loss):
    """'nllNN' -> NN/100 clipped to [0,1]; unreadable -> 1.0 = plain NLL (outlayers.py:20-27)."""
    try:
        return float(np.clip(int(loss[-2:]) / 100, 0, 1))
    except ValueError:
        return 1.0


def output_loss(kind, loss, z, logprob, y, Bg):
    """(cost term, dL/dz) with the mean taken over Bg samples.  maximum(0, a) hands the gradient
    on where a >= 0 (A5)."""
    dt = z.dtype
    one = dt.type(1)
    B, n = z.shape
    rows = np.arange(B)
    onehot = np.zeros_like(z)
    onehot[rows, y] = one
    if kind == 'SoftmaxLayer':
        lpy = logprob[rows, y]
        p = np.exp(logprob)
        if loss == 'nll':                                                    # :50-51
            per, dl = -lpy, -np.ones_like(lpy)
        elif loss == 'nllsq':                                                # :41-42
            per, dl = lpy * lpy, dt.type(2) * lpy
        elif loss.startswith('nll'):                                         # :44-48
            with np.errstate(divide='ignore'):
                a = dt.type(np.log(loss_threshold(loss))) - lpy
            per, dl = np.maximum(dt.type(0), a), -(a >= 0).astype(dt)
        else:
            raise NotImplementedError("Loss : " + loss)
        g = dl[:, None] * (onehot - p)
    elif kind == 'ExpLossLayer':                                             # :38-39
        o = z - z.mean(axis=1, keepdims=True, dtype=dt)
        per = np.exp(-o[rows, y])
        g = -per[:, None] * (onehot - one / dt.type(n))
    elif kind == 'HingeLayer':                                               # :62-64, mean over B*n
        a = z + one - z[rows, y][:, None]
        m = (a >= 0).astype(dt)
        per = (np.maximum(dt.type(0), a)).sum(axis=1, dtype=dt) / dt.type(n)
        g = (m - onehot * m.sum(axis=1, keepdims=True, dtype=dt)) / dt.type(n)
    else:
        raise NotImplementedError(kind)
    return per.sum(dtype=dt) / dt.type(Bg), (g / dt.type(Bg)).astype(dt)


# ----------------------------------------------------------------------------------------------
# The network -- theanet/neuralnet.py:59-333
# ----------------------------------------------------------------------------------------------
def bf16_round(a):
    """float32 -> bfloat16 (round to nearest even) -> float32, the rounding the tensor-core conv
    path applies to its operands and outputs (theanet_b200/csrc/conv_tc.cu)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    r = (u + np.uint32(0x7fff) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xffff0000)
    return r.view(np.float32).reshape(np.shape(a))


def conv_tc_eligible(in_maps, in_sz, M, f, mode, actvn, out_sz, pool_sz=None, ignore_border=False,
                     first=False):
    """Which ConvLayers theanet_b200 runs as bf16 tensor-core implicit GEMMs when
    training_params['CONV_DTYPE'] == 'bfloat16' (mirrors NeuralNet._conv_tc_kind): wide layers
    directly, the first weighted layer through a 64-wide im2col when C*f*f <= 64."""
    fast = actvn in ('linear', 'relu') or (len(actvn) == 6 and actvn.startswith('relu'))
    wide = in_maps % 64 == 0 or (first and f == 3 and in_maps <= 7)
    if not (wide and M % 64 == 0 and mode == 'same' and fast and 1 <= f <= 7):
        return False
    if out_sz < 4 or out_sz > 128 or out_sz & (out_sz - 1):
        return False
    if pool_sz is not None and (pool_sz != 2 or out_sz % 2 or ignore_border):
        return False
    return True


class OracleNet:
    """Eager numpy twin of NeuralNet: same constructor arguments, same RNG consumption order for
    initialisation, same train/test semantics.

    train_step(x, y, step, sample0) consumes an already sliced minibatch (the reference slices a
    resident corpus by batch index, neuralnet.py:219-226; slicing is the caller's job here).
    """

    def __init__(self, layers, training_params, allwts=None, dtype=np.float32, img_sz=None):
        self.dtype = np.dtype(dtype)
        self.layers = layers
        self.tr_prms = training_params
        rand_gen = np.random.RandomState(training_params['SEED']) if allwts is None else None
        self.batch_sz = training_params['BATCH_SZ']
        self.random = PhiloxRandom()
        self.tie_source = None  # {PoolLayer index: activations of another implementation}, see train_step
        self.kink_source = None  # {HiddenLayer index: output of another implementation}, see train_step
        self.kink_flips = []     # (layer, elements re-signed, max |z| among them, max |z|) per step
        self.spec = []          # per layer dict: kind, args, params, vel, seeds, shapes
        num_maps, out_sz, n_out = None, None, None
        for li, (name, args) in enumerate(layers):
            args = dict(args)
            wts = allwts[li] if allwts else None
            L = {'kind': name, 'args': args, 'params': [], 'vel': None, 'reg': None}
            if name in ('InputLayer', 'ElasticLayer'):
                if li > 0:
                    args['num_maps'], args['img_sz'] = num_maps, out_sz     # neuralnet.py:132-142
                elif img_sz is not None and 'img_sz' not in args:
                    args['img_sz'] = img_sz
                num_maps = args.get('num_maps', 1)
                out_sz = args['img_sz']
                n_out = num_maps * out_sz ** 2
                if name == 'ElasticLayer':
                    assert args.get('zoom', 1) > 0
                    L['identity'] = elastic_is_identity(args)
                    if not L['identity']:                                    # inlayers.py:72
                        L['seed'] = int(rand_gen.randint(1e6)) if rand_gen is not None \
                            else int(np.random.randint(1e6))
                L['num_maps'], L['out_sz'] = num_maps, out_sz
            elif name == 'ConvLayer':
                f, M = args['filter_sz'], args['num_maps']
                stride = args.get('stride', 1)
                actvn = args.get('actvn', 'relu50')
                mode = args.get('mode', 'valid')
                assert stride == 1 or mode == 'valid', "For Same mode stride should be 1"  # convpool.py:58
                if wts is None:
                    W, b = init_wb(rand_gen, (M, num_maps, f, f), (M,),
                                   num_maps * f * f, M * f * f, actvn)       # convpool.py:46-50
                else:
                    W, b = [np.asarray(t, np.float32) for t in wts]
                _, out_sz = conv_geometry(out_sz, f, mode)
                # conv2d(subsample=(s,s)) samples the stride-1 output at (i*s, j*s); the reference
                # books out_sz // s (convpool.py:70), which is that size only when s divides it
                assert out_sz % stride == 0, "out_sz is not a multiple of the stride"
                out_sz //= stride
                L.update(actvn=actvn, mode=mode, in_maps=num_maps, stride=stride)
                num_maps = M
                n_out = M * out_sz ** 2
                L['params'] = [W.astype(self.dtype), b.astype(self.dtype)]
                L['reg'] = dict(DEFAULT_REG, **dict(args.get('reg', ())))
                L['num_maps'], L['out_sz'] = num_maps, out_sz
            elif name == 'PoolLayer':
                out_sz = pool_out_size(out_sz, args['pool_sz'], args.get('ignore_border', False))
                n_out = num_maps * out_sz ** 2
                L['num_maps'], L['out_sz'] = num_maps, out_sz
            elif name == 'MeanLayer':                                        # convpool.py:129-144
                L['in_sz'] = out_sz
                out_sz, n_out = 1, num_maps
                L['num_maps'], L['out_sz'] = num_maps, out_sz
            elif name == 'ColorLayer':                                       # color.py:9-52
                if li > 0:
                    args['num_maps'], args['img_sz'] = num_maps, out_sz     # neuralnet.py:132-142
                num_maps = args.get('num_maps', 3)
                out_sz = args['img_sz']
                n_out = num_maps * out_sz ** 2
                L['identity'] = args.get('gamma', 1) == 1 and args.get('balance', 1) == 1
                if not L['identity']:
                    assert args.get('gamma', 1) > 0 and args.get('balance', 1) > 0
                    L['seed'] = int(rand_gen.randint(1e6)) if rand_gen is not None \
                        else int(np.random.randint(1e6))                     # color.py:30
                L['num_maps'], L['out_sz'] = num_maps, out_sz
            elif name == 'DropOutLayer':
                if args.get('pdrop', 0):                                     # dropout.py:10
                    L['seed'] = int(rand_gen.randint(1e6)) if rand_gen is not None \
                        else int(np.random.randint(1e6))
            elif name in ('HiddenLayer',) + OUT_LAYERS:
                n_in = n_out
                n_o = args['n_out']
                if name == 'HiddenLayer':
                    actvn = init_act = args.get('actvn', 'relu01')
                elif name == 'SoftmaxLayer':
                    actvn, init_act = 'softmax', 'Softmax'                   # outlayers.py:88
                    L['loss'] = args.get('loss', 'nll')
                else:                                                        # outlayers.py:108,132
                    actvn = init_act = 'linear'
                    L['loss'] = 'exp' if name == 'ExpLossLayer' else 'hinge' 
                if wts is None:
                    W, b = init_wb(rand_gen, (n_in, n_o), (n_o,), n_in + n_o, n_in + n_o,
                                   init_act)                                 # hidden.py:21-27
                else:
                    W, b = [np.asarray(t, np.float32) for t in wts]
                pdrop = args.get('pdrop', 0) if name == 'HiddenLayer' else 0
                assert name == 'HiddenLayer' or li == len(layers) - 1, "output layer must be last"
                if pdrop:                                                    # dropout.py:10
                    L['seed'] = int(rand_gen.randint(1e6)) if rand_gen is not None \
                        else int(np.random.randint(1e6))
                L.update(actvn=actvn, pdrop=pdrop)
                L['params'] = [W.astype(self.dtype), b.astype(self.dtype)]
                L['reg'] = dict(DEFAULT_REG, **dict(args.get('reg', ())))
                n_out = n_o
                num_maps, out_sz = None, None
            elif name in ('AuxConcatLayer', 'SoftAuxLayer'):                 # auxiliary.py:60-160
                n_in = n_out
                nh, no = args['n_aux']
                assert args['aux_type'] == 'LocationInfo'
                ws = None if wts is None else [np.asarray(t, np.float32) for t in wts]
                hidden = []
                if name == 'SoftAuxLayer':                                   # :108-111
                    n_o = args['n_out']
                    hidden = list(init_wb(rand_gen, (n_in, n_o), (n_o,), n_in + n_o, n_in + n_o,
                                          'linear')) if ws is None else ws[:2]
                    aux_ws = None if ws is None else ws[2:6]
                else:
                    aux_ws = ws
                if rand_gen is not None:                                     # :25 (stream first)
                    L['seed'] = int(rand_gen.randint(1e6))
                else:
                    L['seed'] = int(np.random.randint(1e6))
                if aux_ws is None:                                           # :36-53
                    aux_ws = list(init_wb(rand_gen, (2, nh), nh, nh + 2, nh + 2, 'relu50')) + \
                        list(init_wb(rand_gen, (nh, no), no, no + nh, no + nh, 'relu01'))
                cross = []
                if name == 'SoftAuxLayer':                                   # :121-126
                    cross = list(init_wb(rand_gen, (no, n_o), n_o, no + n_o, no + n_o,
                                         'softmax')) if ws is None else ws[6:]
                    L['reg'] = dict(DEFAULT_REG, **dict(args.get('reg', ())))
                    L['loss'] = args.get('loss', 'nll')
                    n_out = n_o
                    assert li == len(layers) - 1, "output layer must be last"
                else:                                                        # no reg: never updated
                    n_out = n_in + no
                L.update(pdrop=0, boost=args.get('boost', 1), n_in=n_in)
                L['params'] = [t.astype(self.dtype) for t in hidden + aux_ws + cross]
                num_maps, out_sz = None, None
            else:
                raise NotImplementedError("Unknown Layer Type" + name)
            L['n_out'] = n_out
            if L['params']:
                L['vel'] = [np.zeros_like(p) for p in L['params']]
            self.spec.append(L)
        # NOT reference behaviour: theanet_b200's optional mixed-precision conv stack (config C4).
        # The restatement rounds the same tensors to bfloat16 at the same points so that the
        # tensor-core path can be checked tightly (float32 accumulation on both sides).
        bf16 = str(training_params.get('CONV_DTYPE', 'float32')).lower() in ('bf16', 'bfloat16')
        for li, L in enumerate(self.spec):
            L['tc'] = False
            if bf16 and L['kind'] == 'ConvLayer':
                nxt = self.spec[li + 1] if li + 1 < len(self.spec) else None
                if nxt is not None and nxt['kind'] == 'MeanLayer':
                    continue                                # stays float32 in the product as well
                pool = nxt['args'] if nxt is not None and nxt['kind'] == 'PoolLayer' else None
                in_sz = self.spec[li - 1].get('out_sz')
                L['tc'] = conv_tc_eligible(
                    L['in_maps'], in_sz, L['num_maps'], L['args']['filter_sz'], L['mode'], L['actvn'],
                    L['out_sz'], pool['pool_sz'] if pool else None,
                    pool.get('ignore_border', False) if pool else False,
                    first=not any(P['params'] for P in self.spec[:li]))
        if 'CUR_EPOCH' not in training_params:                               # neuralnet.py:108-109
            training_params['CUR_EPOCH'] = 0
        self.set_rate()
        self.injected = None

    # -- learning-rate schedule (neuralnet.py:303-314) -----------------------------------------
    def set_rate(self):
        self.cur_learn_rate = np.float32(
            self.tr_prms['INIT_LEARNING_RATE'] /
            (1 + self.tr_prms['CUR_EPOCH'] / self.tr_prms['EPOCHS_TO_HALF_RATE']))

    def inc_epoch_set_rate(self):
        self.tr_prms['CUR_EPOCH'] += 1
        self.set_rate()

    def get_wts(self):
        return [[p.astype(np.float32) for p in L['params']] for L in self.spec]

    # -- forward --------------------------------------------------------------------------------
    def _forward(self, x, train, step, sample0, rand=None, aux=None):
        dt = self.dtype
        a = np.asarray(x, dt)
        B = a.shape[0]
        samples = np.arange(sample0, sample0 + B)
        caches = []
        for li, L in enumerate(self.spec):
            kind, args = L['kind'], L['args']
            c = {}
            if kind == 'InputLayer':
                pass
            elif kind == 'ElasticLayer':
                if train and not L['identity']:
                    h = L['out_sz']
                    if rand is not None and (li, 'noise') in rand:
                        noise, u = rand[(li, 'noise')], rand[(li, 'u')]
                    else:
                        noise, u = self.random.elastic(L['seed'], step, h)
                    ty, tx, disp = elastic_target(h, args, noise, u)
                    fm = None
                    if args.get('pflip', 0):
                        n = a[0].size
                        if rand is not None and (li, 'flip') in rand:
                            fm = rand[(li, 'flip')]
                        else:
                            fm = self.random.flip_mask(L['seed'], step, samples, n, args['pflip'])
                        fm = fm.reshape(a.shape)
                    a = elastic_apply(a, args, ty, tx, fm)
                    c['disp'] = disp
                else:                                                        # inlayers.py:67-70,157-163
                    a = elastic_apply(a, args, None, None, None)
            elif kind == 'ConvLayer':
                if a.ndim != 4:
                    raise ValueError("conv after a dense layer")
                W, b = L['params']
                if L['tc']:                                  # bf16 operands, float32 accumulate
                    a, W = bf16_round(a).astype(dt), bf16_round(W).astype(dt)
                    c['Wr'] = W
                z, cc = conv_forward(a, W, L['mode'])
                if L['stride'] != 1:
                    c['full_shape'] = z.shape
                    z = np.ascontiguousarray(z[:, :, ::L['stride'], ::L['stride']])
                z = z + b[None, :, None, None]
                c.update(cc=cc, z=z)
                a = act_forward(L['actvn'], z)
                if L['tc']:
                    a = bf16_round(a).astype(dt)             # activations are stored in bf16
                c['a'] = a
            elif kind == 'PoolLayer':
                a, pc = pool_forward(a, args['pool_sz'], args.get('ignore_border', False))
                c['pc'] = pc
            elif kind == 'MeanLayer':
                c['in_shape'] = a.shape
                a = a.sum(axis=(2, 3), dtype=dt) / dt.type(a.shape[2] * a.shape[3])
            elif kind == 'ColorLayer':
                if train and not L['identity']:
                    if rand is not None and (li, 'color') in rand:
                        u = np.asarray(rand[(li, 'color')], np.float32)
                    else:
                        u = philox.color_uniforms(L['seed'], step, samples, a.shape[1])
                    a = color_jitter(a, args, u.reshape(B, a.shape[1], 3))
            elif kind == 'DropOutLayer':
                p = args.get('pdrop', 0)
                if p:
                    if train:
                        n = a[0].size
                        if rand is not None and (li, 'mask') in rand:
                            m = rand[(li, 'mask')]
                        else:
                            m = self.random.keep_mask(L['seed'], step, samples, n, 1 - p)
                        m = m.reshape(a.shape).astype(dt)
                        a = a * m
                        c['mask'] = m
                    else:                                                    # dropout.py:28-31
                        a = a * dt.type(1 - p)
            elif kind in ('AuxConcatLayer', 'SoftAuxLayer'):
                assert aux is not None, "Auxillary data not supplied"
                xin = a.reshape(B, -1)
                P = L['params']
                k0 = 2 if kind == 'SoftAuxLayer' else 0
                w1, b1, w2, b2 = P[k0:k0 + 4]
                A = np.asarray(aux, dt).reshape(B, 2, 2)
                if train:                                                    # auxiliary.py:25-28
                    if rand is not None and (li, 'auxu') in rand:
                        u = np.asarray(rand[(li, 'auxu')], np.float32)
                    else:
                        u = philox.aux_uniforms(L['seed'], step, samples)
                    u = u.astype(dt).reshape(B, 1)
                    loc = A[:, 0, :] * u + A[:, 1, :] * (dt.type(1) - u)
                else:                                                        # :31
                    loc = (A[:, 0, :] + A[:, 1, :]) / dt.type(2)
                loc = loc * dt.type(L['boost'])                              # :33
                z1 = loc @ w1 + b1
                hid = act_forward('relu50', z1)
                z2 = hid @ w2 + b2
                feat = act_forward('relu01', z2)
                c.update(x=xin, in_shape=a.shape, loc=loc, z1=z1, hid=hid, z2=z2, aux=feat)
                if kind == 'AuxConcatLayer':                                 # :80
                    a = np.concatenate([xin, feat], axis=1)
                else:                                                        # :133-134
                    W, b, cw, cb = P[0], P[1], P[6], P[7]
                    z = ((xin @ W + b) + cb) + feat @ cw
                    c['z'] = z
                    c['features'], a, c['probs'] = output_views('SoftmaxLayer', z)
            elif kind in ('HiddenLayer',) + OUT_LAYERS:
                xin = a.reshape(B, -1)                                       # neuralnet.py:168-169
                W, b = L['params']
                z = xin @ W + b
                c.update(x=xin, z=z, in_shape=a.shape)
                if kind in OUT_LAYERS:
                    c['features'], a, c['probs'] = output_views(kind, z)     # a = the layer's logprob
                else:
                    a = act_forward(L['actvn'], z)
                    c['a'] = a
                    p = L['pdrop']
                    if p:
                        if train:
                            if rand is not None and (li, 'mask') in rand:
                                m = rand[(li, 'mask')]
                            else:
                                m = self.random.keep_mask(L['seed'], step, samples, a.shape[1], 1 - p)
                            m = m.reshape(a.shape).astype(dt)
                            a = a * m                                        # dropout.py:9-13, no rescale
                            c['mask'] = m
                        else:
                            a = a * dt.type(1 - p)                           # hidden.py:50-55
            c['out'] = a
            caches.append(c)
        return a, caches

    # -- one training step (neuralnet.py:203-241, layer.py:70-117) --------------------------------
    def train_step(self, x, y, step=0, sample0=0, rand=None, global_batch=None, apply_update=True,
                   aux=None):
        """Returns (cost, logprob).  ``global_batch`` (default: len(x)) is the divisor of the mean
        NLL so that a data-parallel shard contributes its share of the global-batch gradient."""
        dt = self.dtype
        y = np.asarray(y, np.int64)
        logprob, caches = self._forward(x, True, step, sample0, rand, aux)
        B = logprob.shape[0]
        Bg = B if global_batch is None else global_batch
        top = self.spec[-1]
        top_kind = 'SoftmaxLayer' if top['kind'] == 'SoftAuxLayer' else top['kind']
        nll, g = output_loss(top_kind, top['loss'], caches[-1]['z'], logprob, y, Bg)
        self.last_features = caches[-1]['features']
        wtcost = dt.type(0)
        for L in self.spec:                                                  # layer.py:109-117
            if L['reg'] is not None:
                l1, l2 = dt.type(L['reg']['L1']), dt.type(L['reg']['L2'])
                if l1:
                    wtcost += l1 * sum(np.abs(t).sum(dtype=dt) for t in L['params'])
                if l2:
                    wtcost += l2 * sum((t * t).sum(dtype=dt) for t in L['params'])
        cost = nll + wtcost
        # backward: g = dL/dz of the output layer
        grads = [None] * len(self.spec)
        first_weighted = min(i for i, L in enumerate(self.spec) if L['params'])
        for li in range(len(self.spec) - 1, -1, -1):
            L, c = self.spec[li], caches[li]
            kind = L['kind']
            if kind == 'SoftAuxLayer':
                P = L['params']
                W, w2, cw = P[0], P[4], P[6]
                gz2 = act_backward('relu01', c['z2'], c['aux'], g @ cw.T)
                gz1 = act_backward('relu50', c['z1'], c['hid'], gz2 @ w2.T)
                grads[li] = [c['x'].T @ g, g.sum(axis=0),
                             c['loc'].T @ gz1, gz1.sum(axis=0), c['hid'].T @ gz2, gz2.sum(axis=0),
                             c['aux'].T @ g, g.sum(axis=0)]
                g = (g @ W.T).reshape(c['in_shape']) if li > first_weighted else None
            elif kind == 'AuxConcatLayer':
                if g is not None:
                    g = np.ascontiguousarray(g[:, :L['n_in']]).reshape(c['in_shape'])
            elif kind in OUT_LAYERS:
                W, b = L['params']
                grads[li] = [c['x'].T @ g, g.sum(axis=0)]
                g = (g @ W.T).reshape(c['in_shape']) if li > first_weighted else None
            elif kind == 'HiddenLayer':
                W, b = L['params']
                if 'mask' in c:
                    g = g * c['mask']
                z = c['z']
                if self.kink_source and li in self.kink_source:
                    # kink localisation (tests): the derivative of the ReLU family jumps at z = 0, and
                    # whether a pre-activation of ~1e-7 lands left or right of it depends on the
                    # summation order of the product in front.  Take the side the OTHER implementation
                    # landed on (read off its output where the dropout mask kept it) and record how
                    # many elements that touched and how close to zero they were.
                    other = np.asarray(self.kink_source[li], dt).reshape(z.shape)
                    kept = c['mask'] != 0 if 'mask' in c else np.ones(z.shape, bool)
                    flip = kept & (other != 0) & (z != 0) & ((other > 0) != (z > 0))
                    self.kink_flips.append((li, int(flip.sum()), float(np.abs(z[flip]).max()) if flip.any() else 0.,
                                            float(np.abs(z).max())))
                    z = np.where(flip, np.copysign(z, other), z)
                gz = act_backward(L['actvn'], z, c['a'], g)
                grads[li] = [c['x'].T @ gz, gz.sum(axis=0)]
                g = (gz @ W.T).reshape(c['in_shape']) if li > first_weighted else None
            elif kind == 'DropOutLayer':
                if g is not None and 'mask' in c:
                    g = g * c['mask']
            elif kind == 'PoolLayer':
                if g is not None:
                    pc = c['pc']
                    if self.tie_source and li in self.tie_source:
                        # tie localisation (tests): route the gradient by the tie pattern of ANOTHER
                        # implementation's activations (same values up to summation order; which
                        # mathematically equal conv sums stay equal in float32 depends on that order)
                        _, pc = pool_forward(np.asarray(self.tie_source[li], dt), L['args']['pool_sz'],
                                             L['args'].get('ignore_border', False))
                    g = pool_backward(g, pc)
            elif kind == 'MeanLayer':
                if g is not None:
                    _, _, h_, w_ = c['in_shape']
                    g = np.broadcast_to((g / dt.type(h_ * w_))[:, :, None, None], c['in_shape']).astype(dt)
            elif kind == 'ColorLayer':
                assert li <= first_weighted, "no gradient through ColorLayer is implemented"
                g = None
            elif kind == 'ConvLayer':
                W, b = L['params']
                gz = act_backward(L['actvn'], c['z'], c['a'], g)
                if L['tc']:
                    gz, W = bf16_round(gz).astype(dt), c['Wr']
                if L['stride'] != 1:
                    gfull = np.zeros(c['full_shape'], gz.dtype)
                    gfull[:, :, ::L['stride'], ::L['stride']] = gz
                    gz = gfull
                dW, db, dx = conv_backward(gz, W, c['cc'], need_dx=li > first_weighted)
                if L['tc'] and dx is not None:
                    dx = bf16_round(dx).astype(dt)
                grads[li] = [dW, db]
                g = dx
            else:
                g = None        # input layers: no gradient wrt data
        self.last_grads = grads
        self.last_caches = caches
        if apply_update:
            self.apply_update(grads)
        return cost, logprob

    def apply_update(self, grads):
        dt = self.dtype
        for li, L in enumerate(self.spec):
            if L['reg'] is None or not L['reg']['rate'] or grads[li] is None:    # layer.py:74-75
                continue
            reg = L['reg']
            for k in range(len(L['params'])):
                th = L['params'][k]
                gk = grads[li][k].astype(dt)
                if reg['L1']:
                    gk = gk + dt.type(reg['L1']) * np.sign(th)
                if reg['L2']:
                    gk = gk + dt.type(2 * reg['L2']) * th
                L['params'][k], L['vel'][k] = sgd_update(th, L['vel'][k], gk, reg, self.cur_learn_rate)

    # -- test twin (neuralnet.py:257-296, outlayers.py:66-80) ---------------------------------------
    def test_step(self, x, y, aux=None):
        """(mean(pred != y), mean(probs[y]), logprob, y_preds): sym_and_oth_err_rate for the
        non-LOGIT kinds; probs are the raw scores for HingeLayer (outlayers.py:138)."""
        logprob, caches = self._forward(x, False, 0, 0, None, aux)
        y = np.asarray(y, np.int64)
        B = logprob.shape[0]
        probs = caches[-1]['probs']
        self.last_features = caches[-1]['features']
        y_preds = np.argmax(caches[-1]['z'], axis=1)          # argmax of scores = of probs / of o
        return (np.mean(y_preds != y).astype(self.dtype),
                np.mean(probs[np.arange(B), y], dtype=self.dtype), logprob, y_preds)
