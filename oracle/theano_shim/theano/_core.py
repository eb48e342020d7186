"""Lazy expression graph with Theano's surface, evaluated by torch on the CPU.
TEST INFRASTRUCTURE ONLY -- see ../README.md.

Dtype rules follow Theano with floatX=float32: tensor (+) tensor promotes like NumPy
(int64 (+) float32 -> float64); Python / NumPy *scalars* never widen a float tensor
(theano.scalar NumpyAutocaster casts Python floats to floatX).
"""
import builtins
import numpy as np
import torch
import torch.nn.functional as F


class _Config:
    floatX = 'float32'


config = _Config()

_TORCH = {'float32': torch.float32, 'float64': torch.float64, 'int32': torch.int32,
          'int64': torch.int64, 'int8': torch.int8, 'bool': torch.bool, 'uint8': torch.uint8}
_NP_OF = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32,
          torch.int64: np.int64, torch.int8: np.int8, torch.bool: np.bool_, torch.uint8: np.uint8}


class Env:
    """One evaluation of a compiled function: memo of evaluated nodes, givens, random-draw log."""

    def __init__(self, givens=None):
        self.memo = {}
        self.givens = {id(k): v for k, v in (givens or {}).items()}
        self.draws = []


def _ev(a, env):
    if isinstance(a, Var):
        return a.ev(env)
    if isinstance(a, (tuple, list)):
        return type(a)(_ev(v, env) for v in a)
    if isinstance(a, slice):
        return slice(_idx(_ev(a.start, env)), _idx(_ev(a.stop, env)), _idx(_ev(a.step, env)))
    return a


def _idx(v):
    if torch.is_tensor(v) and v.ndim == 0:
        return int(v)
    return v


def _tensor(v):
    if torch.is_tensor(v):
        return v
    return torch.from_numpy(np.ascontiguousarray(v))


def _is_weak(v):
    return isinstance(v, (bool, int, float, np.generic))


def _coerce(a, b):
    """Bring two evaluated operands to a common dtype under the rules in the module docstring."""
    wa, wb = _is_weak(a), _is_weak(b)
    if wa and wb:
        return torch.tensor(float(a) if isinstance(a, (float, np.floating)) else a), \
            torch.tensor(float(b) if isinstance(b, (float, np.floating)) else b)
    if wa or wb:
        t, s = (_tensor(b), a) if wa else (_tensor(a), b)
        if not t.dtype.is_floating_point and isinstance(s, (float, np.floating)):
            t = t.to(_TORCH[config.floatX])
        s = torch.tensor(s, dtype=t.dtype)
        return (s, t) if wa else (t, s)
    a, b = _tensor(a), _tensor(b)
    if a.dtype != b.dtype:
        dt = np.result_type(_NP_OF[a.dtype], _NP_OF[b.dtype])
        dt = _TORCH[np.dtype(dt).name]
        a, b = a.to(dt), b.to(dt)
    return a, b


class Var:
    __array_ufunc__ = None          # numpy operands defer to our reflected operators
    broadcastable = ()

    def __init__(self, fn, args=(), name=None):
        self.fn, self.args, self.name = fn, args, name

    def ev(self, env):
        k = id(self)
        if k in env.memo:
            return env.memo[k]
        if k in env.givens:
            v = _ev(env.givens[k], env)
            want = getattr(self, 'dtype', None)
            if want is not None and torch.is_tensor(v) and v.dtype != _TORCH[want]:
                raise TypeError("given for {} has dtype {}, wants {}".format(self.name, v.dtype, want))
        else:
            v = self.fn(env, *[_ev(a, env) for a in self.args])
        env.memo[k] = v
        return v

    # -- arithmetic -----------------------------------------------------------------------------
    def _bin(self, other, op, swap=False):
        a, b = (other, self) if swap else (self, other)
        return Var(lambda env, x, y: op(*_coerce(x, y)), (a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, _true_div)
    def __rtruediv__(self, o): return self._bin(o, _true_div, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __ge__(self, o): return self._bin(o, torch.ge)
    def __neg__(self): return Var(lambda env, x: -x, (self,))
    def __abs__(self): return Var(lambda env, x: torch.abs(x), (self,))
    __hash__ = object.__hash__

    # -- shape ----------------------------------------------------------------------------------
    @property
    def shape(self):
        return Var(lambda env, x: tuple(x.shape), (self,))

    @property
    def ndim(self):
        raise AttributeError("theano_shim: symbolic ndim is not tracked")

    def __getitem__(self, key):
        def fix(v):
            if torch.is_tensor(v):
                return v if v.dtype == torch.bool else v.long()
            return _idx(v)

        def fn(env, x, k):
            if isinstance(x, tuple):                     # indexing a shape
                return x[_idx(k)]
            return x[tuple(fix(v) for v in k) if isinstance(k, tuple) else fix(k)]
        return Var(fn, (self, key))

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (tuple, list)):
            pattern = tuple(pattern[0])

        def fn(env, x):
            x = x.permute(*[p for p in pattern if p != 'x'])
            for i, p in enumerate(pattern):
                if p == 'x':
                    x = x.unsqueeze(i)
            return x
        return Var(fn, (self,))

    def flatten(self, ndim=1):
        return Var(lambda env, x: x.reshape(tuple(x.shape[:ndim - 1]) + (-1,)), (self,))

    def reshape(self, shape, ndim=None):
        return Var(lambda env, x, s: x.reshape(tuple(_idx(v) for v in s)), (self, tuple(shape)))

    def sum(self, axis=None): return sum(self, axis)
    def mean(self, axis=None): return mean(self, axis)
    def max(self, axis=None): return max(self, axis)
    def astype(self, dtype): return cast(self, dtype)

    def __iter__(self):
        raise TypeError("theano_shim: symbolic variables are not iterable")

    def __repr__(self):
        return "<shim Var {}>".format(self.name or hex(id(self)))


def _true_div(a, b):
    if not a.dtype.is_floating_point and not b.dtype.is_floating_point:
        a, b = a.double(), b.double()          # Theano: int / int -> float64
    return torch.div(a, b)


# ---- leaves -------------------------------------------------------------------------------------
class _Placeholder(Var):
    def __init__(self, name, dtype, ndim):
        Var.__init__(self, self._unbound, (), name)
        self.dtype, self._ndim = dtype, ndim

    def _unbound(self, env):
        raise ValueError("theano_shim: input '{}' has no value (not an input, not in givens)".format(self.name))

    def bind(self, env, value):
        v = torch.as_tensor(np.asarray(value, dtype=self.dtype))
        if v.ndim != self._ndim:
            raise TypeError("input '{}' wants {} dimensions, got {}".format(self.name, self._ndim, v.ndim))
        env.memo[id(self)] = v


def _ph(dtype, ndim):
    return lambda name=None: _Placeholder(name, dtype, ndim)


tensor4, tensor3, matrix, vector, scalar = (_ph('float32', n) for n in (4, 3, 2, 1, 0))
ivector, lvector = _ph('int32', 1), _ph('int64', 1)
iscalar, lscalar = _ph('int32', 0), _ph('int64', 0)


class SharedVar(Var):
    """theano.shared: a numpy array that graphs read and `updates` overwrite.  The class name
    contains 'SharedVar...' on purpose: the reference tests `'SharedVariable' in str(type(x))`
    (theanet/layer/weights.py:13-18) through the alias in theano/compile.py."""

    def __init__(self, value, name=None, borrow=False, broadcastable=None):
        Var.__init__(self, self._read, (), name)
        self.value = np.array(value)
        self.broadcastable = tuple(False for _ in self.value.shape)

    def _read(self, env):
        t = torch.from_numpy(self.value.copy())
        if t.dtype.is_floating_point:
            t.requires_grad_(True)
        return t

    def get_value(self, borrow=False):
        return self.value if borrow else self.value.copy()

    def set_value(self, v, borrow=False):
        self.value = np.asarray(v, dtype=self.value.dtype).reshape(self.value.shape).copy()


SharedVar.__name__ = 'SharedVariable'
SharedVar.__qualname__ = 'SharedVariable'


def shared(value, name=None, borrow=False, broadcastable=None, **kw):
    return SharedVar(value, name, borrow, broadcastable)


def constant(value, name=None):
    if isinstance(value, Var):
        return value
    arr = np.asarray(value)
    t = torch.from_numpy(np.ascontiguousarray(arr))
    return Var(lambda env: t, (), name)


as_tensor_variable = constant


# ---- compiled functions ---------------------------------------------------------------------------
class Function:
    def __init__(self, inputs, outputs, updates, givens):
        self.inputs = list(inputs)
        self.single = isinstance(outputs, Var)
        self.outputs = [] if outputs is None else [outputs] if self.single else list(outputs)
        self.updates = list(updates.items()) if isinstance(updates, dict) else list(updates or ())
        self.givens = dict(givens or {})
        self.draws = []

    def __call__(self, *args):
        if len(args) != len(self.inputs):
            raise TypeError("expected {} inputs".format(len(self.inputs)))
        env = Env(self.givens)
        for v, a in zip(self.inputs, args):
            v.bind(env, a)
        outs = [o.ev(env) for o in self.outputs]
        news = [(s, e.ev(env) if isinstance(e, Var) else e) for s, e in self.updates]
        for s, v in news:                      # all right-hand sides saw the OLD values
            s.set_value(v.detach().numpy() if torch.is_tensor(v) else v)
        self.draws = env.draws
        res = [np.asarray(o.detach().numpy()) if torch.is_tensor(o) else np.asarray(o) for o in outs]
        return res[0] if self.single else res


def function(inputs, outputs=None, updates=None, givens=None, **kw):
    return Function(inputs, outputs, updates, givens)


# ---- elementwise -------------------------------------------------------------------------------------
def _un(op):
    return lambda x: Var(lambda env, v: op(_tensor(v) if not _is_weak(v) else torch.tensor(float(v))), (x,))


exp, log, tanh, cos, sin, sqrt = (_un(f) for f in (torch.exp, torch.log, torch.tanh, torch.cos,
                                                    torch.sin, torch.sqrt))
sqr = _un(lambda v: v * v)
abs_ = _un(torch.abs)
zeros_like = _un(torch.zeros_like)


class _TieMax(torch.autograd.Function):
    """theano.scalar.Maximum / Minimum: out = max(x, y); both inputs receive the gradient where
    they equal the output (grad uses eq(out, x), eq(out, y)) -- assumption A5."""

    @staticmethod
    def forward(ctx, x, y, is_max):
        out = torch.maximum(x, y) if is_max else torch.minimum(x, y)
        ctx.save_for_backward(x, y, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, out = ctx.saved_tensors

        def red(gi, ref):                       # un-broadcast
            while gi.ndim > ref.ndim:
                gi = gi.sum(0)
            for d, n in enumerate(ref.shape):
                if n == 1 and gi.shape[d] != 1:
                    gi = gi.sum(d, keepdim=True)
            return gi
        return red(g * (out == x).to(g.dtype), x), red(g * (out == y).to(g.dtype), y), None


def maximum(a, b):
    return Var(lambda env, x, y: _TieMax.apply(*_coerce(x, y), True), (a, b))


def minimum(a, b):
    return Var(lambda env, x, y: _TieMax.apply(*_coerce(x, y), False), (a, b))


def clip(x, lo, hi):
    def fn(env, v, a, b):
        v, a = _coerce(v, a)
        v, b = _coerce(v, b)
        return torch.minimum(torch.maximum(v, a.to(v.dtype)), b)
    return Var(fn, (x, lo, hi))


def iround(x):
    """theano.tensor.iround: round half away from zero, int64 (A6)."""
    return Var(lambda env, v: (torch.sign(v) * torch.floor(torch.abs(v) + 0.5)).to(torch.int64), (x,))


def cast(x, dtype):
    return Var(lambda env, v: _tensor(v).to(_TORCH[dtype]), (x,))


def neq(a, b):
    return Var(lambda env, x, y: torch.ne(*_coerce(x, y)), (a, b))


def eq(a, b):
    return Var(lambda env, x, y: torch.eq(*_coerce(x, y)), (a, b))


# ---- reductions ------------------------------------------------------------------------------------------
def _axes(axis):
    return None if axis is None else tuple(axis) if isinstance(axis, (tuple, list)) else (axis,)


def sum(x, axis=None):
    ax = _axes(axis)
    return Var(lambda env, v: v.sum() if ax is None else v.sum(dim=ax), (x,))


def mean(x, axis=None):
    ax = _axes(axis)

    def fn(env, v):
        if not v.dtype.is_floating_point:
            v = v.double()                      # Theano: mean of bool/int -> float64
        return v.mean() if ax is None else v.mean(dim=ax)
    return Var(fn, (x,))


def max(x, axis=None):
    return Var(lambda env, v: v.max() if axis is None else v.max(dim=axis).values, (x,))


def argmax(x, axis=None):
    return Var(lambda env, v: v.argmax() if axis is None else v.argmax(dim=axis), (x,))


# ---- construction ----------------------------------------------------------------------------------------------
def arange(n):
    return Var(lambda env, k: torch.arange(_idx(k)), (n,))


def stack(*xs, **kw):
    if len(xs) == 1 and isinstance(xs[0], (tuple, list)):
        xs = tuple(xs[0])
    return Var(lambda env, vs: torch.stack([_tensor(v) for v in vs]), (tuple(xs),))


def concatenate(xs, axis=0):
    return Var(lambda env, vs: torch.cat([_tensor(v) for v in vs], dim=axis), (tuple(xs),))


def dot(a, b):
    return Var(lambda env, x, y: torch.matmul(*_coerce(x, y)), (a, b))


def tensordot(a, b, axes=2):
    def fn(env, x, y):
        x, y = _coerce(x, y)
        ax = axes
        if isinstance(ax, (tuple, list)):
            ax = tuple([v] if isinstance(v, int) else list(v) for v in ax)
        return torch.tensordot(x, y, dims=ax)
    return Var(fn, (a, b))


def grad(cost, wrt):
    def fn(env):
        c, p = cost.ev(env), wrt.ev(env)
        g, = torch.autograd.grad(c, p, retain_graph=True, allow_unused=True)
        return torch.zeros_like(p) if g is None else g
    return Var(fn)


# ---- nnet --------------------------------------------------------------------------------------------------------
class _NamedOp:
    """Callable whose str() is the op name: the reference looks activations up by
    str(tt.nnet.sigmoid) etc. (theanet/layer/layer.py:27-54)."""

    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, x):
        return Var(lambda env, v: self.fn(v), (x,))

    def __str__(self):
        return self.name


sigmoid = _NamedOp('sigmoid', torch.sigmoid)
softplus = _NamedOp('softplus', F.softplus)
softmax = _NamedOp('softmax', lambda v: torch.softmax(v, dim=1))     # row-wise on a matrix


def nnet_conv2d(input, filters, input_shape=None, filter_shape=None, border_mode='valid',
                subsample=(1, 1), filter_flip=True, **kw):
    """theano.tensor.nnet.conv2d: a true convolution (kernels flipped, A1); 'full' pads f-1 (A2)."""
    def fn(env, x, w):
        f = w.shape[-1]
        pad = {'valid': 0, 'full': f - 1, 'half': f // 2}[border_mode]
        if filter_flip:
            w = torch.flip(w, dims=(2, 3))
        # accumulate in float64, round once: identical patches give identical outputs wherever
        # they sit, so pooling ties do not depend on the blocking of the convolution primitive
        return F.conv2d(x.double(), w.double(), stride=tuple(subsample), padding=pad).to(x.dtype)
    return Var(fn, (input, filters))


def signal_conv2d(input, filters, image_shape=None, filter_shape=None, border_mode='valid', **kw):
    """theano.tensor.signal.conv.conv2d: every 2-D slice of `input` convolved with one 2-D filter."""
    def fn(env, x, w):
        x, w = _coerce(x, w)
        pad = w.shape[-1] - 1 if border_mode == 'full' else 0
        lead = x.shape[:-2]
        y = F.conv2d(x.reshape((-1, 1) + tuple(x.shape[-2:])), torch.flip(w, dims=(0, 1))[None, None],
                     padding=pad)
        return y.reshape(tuple(lead) + tuple(y.shape[-2:]))
    return Var(fn, (input, filters))


class _Pool(torch.autograd.Function):
    """pool_2d(mode='max'), window = stride = ds, no padding; out = ceil (ignore_border=False keeps
    partial edge windows) or floor.  MaxPoolGrad hands the output gradient to EVERY element equal to
    its window's maximum (A3)."""

    @staticmethod
    def forward(ctx, x, p, ignore_border):
        H, W = x.shape[-2:]
        if ignore_border:
            oh, ow = H // p, W // p
        else:
            oh, ow = -(-H // p), -(-W // p)
        xp = x.new_full(tuple(x.shape[:-2]) + (oh * p, ow * p), -float('inf'))
        hh, ww = builtins.min(H, oh * p), builtins.min(W, ow * p)
        xp[..., :hh, :ww] = x[..., :hh, :ww]
        win = xp.reshape(tuple(x.shape[:-2]) + (oh, p, ow, p))
        out = win.amax(dim=(-3, -1))
        ctx.save_for_backward(win, out)
        ctx.geom = (H, W, hh, ww, oh, ow, p)
        return out

    @staticmethod
    def backward(ctx, g):
        win, out = ctx.saved_tensors
        H, W, hh, ww, oh, ow, p = ctx.geom
        hit = (win == out[..., :, None, :, None]).to(g.dtype)
        gp = (hit * g[..., :, None, :, None]).reshape(tuple(g.shape[:-2]) + (oh * p, ow * p))
        gx = g.new_zeros(tuple(g.shape[:-2]) + (H, W))
        gx[..., :hh, :ww] = gp[..., :hh, :ww]
        return gx, None, None


def pool_2d(input, ds=None, ignore_border=None, st=None, padding=(0, 0), mode='max', ws=None, **kw):
    ds = ws if ds is None else ds
    if mode != 'max' or tuple(padding) != (0, 0) or st not in (None, tuple(ds)) or ds[0] != ds[1]:
        raise NotImplementedError("theano_shim.pool_2d: only square, unpadded, stride=window max pooling")
    return Var(lambda env, x: _Pool.apply(x, int(ds[0]), bool(ignore_border)), (input,))


# ---- random streams --------------------------------------------------------------------------------------------------
class RandomStreams:
    """theano.tensor.shared_randomstreams.RandomStreams.  The values are NOT Theano's (its
    per-variable MT19937 seeding is not reproduced; the oracle does not depend on it, A8): each
    draw comes from one numpy RandomState per stream and is logged as
    (stream index, variable serial within the stream, kind, value) in Function.draws."""
    instances = []

    def __init__(self, seed=None):
        self.rng = np.random.RandomState(seed)
        self.index = len(RandomStreams.instances)
        self.nvars = 0
        RandomStreams.instances.append(self)

    def _var(self, kind, size, draw):
        serial = self.nvars
        self.nvars += 1

        def fn(env, sz):
            sz = tuple(_idx(v) for v in sz) if isinstance(sz, (tuple, list)) else sz
            val = draw(sz)
            env.draws.append((self.index, serial, kind, val))
            return torch.from_numpy(np.ascontiguousarray(val))
        return Var(fn, (size,))

    def uniform(self, size=(), low=0.0, high=1.0, ndim=None, dtype=None):
        dt = dtype or config.floatX
        return self._var('uniform', size, lambda sz: np.asarray(self.rng.uniform(low, high, sz), dt))

    def normal(self, size=(), avg=0.0, std=1.0, ndim=None, dtype=None):
        dt = dtype or config.floatX
        return self._var('normal', size, lambda sz: np.asarray(self.rng.normal(avg, std, sz), dt))

    def binomial(self, size=(), n=1, p=0.5, ndim=None, dtype='int64'):
        return self._var('binomial', size, lambda sz: np.asarray(self.rng.binomial(n, p, sz), dtype))
