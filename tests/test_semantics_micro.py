"""Hand-derived micro-vectors for the Theano op semantics the parity chain rests on (SURVEY.md 8c,
assumptions A1-A7).  Theano itself cannot run here, so the oracle and the stand-in
(oracle/theano_shim) were both written from its documented behaviour; what keeps them honest is
that every expected number below is worked out BY HAND in the comment next to it (no call into any
of the implementations under test), for inputs small enough to check on paper -- including the
(5,5), ds=(2,2) -> (3,3) pooling case the reference itself quotes (theanet/layer/convpool.py:102-103).
Each vector is checked against the shim, the oracle (CPU tests) and the CUDA kernels through the
C ABI (-m gpu)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import theanet_oracle as O   # noqa: E402

F32 = np.float32

# ---- A1: nnet.conv2d is a TRUE convolution (filter_flip=True), 'valid' -----------------------------
# x = 0..15 as 4x4, W (3x3) has W[0,0] = 1 and W[2,2] = 2, zeros elsewhere.
# out[i,j] = sum_{u,v} x[i+u, j+v] * W[2-u, 2-v]: W[0,0] pairs with (u,v) = (2,2), W[2,2] with (0,0):
# out[i,j] = x[i+2, j+2] + 2 x[i, j]
#   out[0,0] = 10 + 0 = 10    out[0,1] = 11 + 2 = 13    out[1,0] = 14 + 8 = 22    out[1,1] = 15 + 10 = 25
# (a correlation would give 2 x[i+2,j+2] + x[i,j] = 20, 23, 32, 35)
A1_X = np.arange(16, dtype=F32).reshape(1, 1, 4, 4)
A1_W = np.zeros((1, 1, 3, 3), F32)
A1_W[0, 0, 0, 0], A1_W[0, 0, 2, 2] = 1, 2
A1_OUT = np.array([[10, 13], [22, 25]], F32).reshape(1, 1, 2, 2)
# gradient wrt W of L = sum(out * G), G = [[1, 0], [0, 0]] (only out[0,0] counts):
# out[0,0] = sum x[u,v] W[2-u,2-v]  =>  dL/dW[a,b] = x[2-a, 2-b]:
#   dW = [[x22, x21, x20], [x12, x11, x10], [x02, x01, x00]] = [[10, 9, 8], [6, 5, 4], [2, 1, 0]]
A1_G = np.array([[1, 0], [0, 0]], F32).reshape(1, 1, 2, 2)
A1_DW = np.array([[10, 9, 8], [6, 5, 4], [2, 1, 0]], F32).reshape(1, 1, 3, 3)

# ---- A2: pool_2d(ds=(2,2)), ignore_border=False keeps the partial edge windows: (5,5) -> (3,3) ------
# x = 0..24 as 5x5; windows = rows {0,1},{2,3},{4} x cols {0,1},{2,3},{4}; the max of a window of an
# increasing ramp is its bottom-right element:
#   [[x11, x13, x14], [x31, x33, x34], [x41, x43, x44]] = [[6, 8, 9], [16, 18, 19], [21, 23, 24]]
# ignore_border=True drops the partial windows: [[6, 8], [16, 18]]
A2_X = np.arange(25, dtype=F32).reshape(1, 1, 5, 5)
A2_OUT = np.array([[6, 8, 9], [16, 18, 19], [21, 23, 24]], F32).reshape(1, 1, 3, 3)
A2_OUT_IB = np.array([[6, 8], [16, 18]], F32).reshape(1, 1, 2, 2)

# ---- A3: MaxPoolGrad gives the output gradient to EVERY element equal to the window's maximum -------
# x (3x3, ceil mode -> 2x2 windows {0,1}x{0,1}, {0,1}x{2}, {2}x{0,1}, {2}x{2}):
#   [[1, 1, 7],     window (0,0) = {1,1,1,0}: max 1, three ties      dout = [[5, 2],
#    [1, 0, 7],     window (0,1) = {7,7}:     max 7, two ties                [3, 4]]
#    [4, 4, 9]]     window (1,0) = {4,4}: two ties; window (1,1) = {9}
#   dx = [[5, 5, 2], [5, 0, 2], [3, 3, 4]]
A3_X = np.array([[1, 1, 7], [1, 0, 7], [4, 4, 9]], F32).reshape(1, 1, 3, 3)
A3_DOUT = np.array([[5, 2], [3, 4]], F32).reshape(1, 1, 2, 2)
A3_DX = np.array([[5, 5, 2], [5, 0, 2], [3, 3, 4]], F32).reshape(1, 1, 3, 3)

# ---- A4: stable log-softmax; argmax = FIRST maximum ------------------------------------------------
# z = [1000, 1000, 999]: log sum exp = 1000 + log(1 + 1 + e^-1) = 1000 + log(2.36787944...) = 1000.86199...
#   logp = [-0.861994..., -0.861994..., -1.861994...]   (a naive exp(1000) overflows)
# z = [3, 3, 1] has two maxima: prediction = index 0
A4_Z = np.array([[1000, 1000, 999], [3, 3, 1]], F32)
A4_LOGP0 = np.array([-0.8619948, -0.8619948, -1.8619948], F32)
A4_PRED = np.array([0, 0])

# ---- A5: reluNN = max(0,x) + min(0,x)*NN/100; at exactly 0 both the maximum and the minimum gradient
# fire (eq(out, arg)): slope 1 + NN/100.  relu10: z = [-2, 0, 3] -> a = [-0.2, 0, 3], slopes [0.1, 1.1, 1]
A5_Z = np.array([-2, 0, 3], F32)
A5_A = np.array([-0.2, 0, 3], F32)
A5_SLOPE = np.array([0.1, 1.1, 1.0], F32)

# ---- A6: tt.iround rounds half AWAY from zero; cast(., 'int32') truncates toward zero ----------------
A6_IN = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, -2.5], np.float64)
A6_ROUND = np.array([1, 2, 3, -1, -2, 2, -3])

# ---- A7: simultaneous updates -> lagged momentum (layer.py:82-86) -----------------------------------
# theta = 1, v = 0.5, g = 2, momentum 0.9, rate 1, lr 0.1:
#   v' = 0.9*0.5 + 0.1*2 = 0.65       theta' = 1 - 0.1 * 0.5 (the OLD v) = 0.95   (not 1 - 0.065)
A7 = dict(theta=1.0, v=0.5, g=2.0, m=0.9, lr=0.1, v_new=0.65, theta_new=0.95)


def close(a, b, tol=1e-6):
    return np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=tol, atol=tol)


# --------------------------------------------------------------------------------------------
# the oracle
# --------------------------------------------------------------------------------------------
def test_oracle_against_hand_derived_vectors():
    z, cache = O.conv_forward(A1_X, A1_W, 'valid')
    assert np.array_equal(z, A1_OUT)
    dW, _, _ = O.conv_backward(A1_G, A1_W, cache, need_dx=False)
    assert np.array_equal(dW, A1_DW)
    out, _ = O.pool_forward(A2_X, 2, False)
    assert np.array_equal(out, A2_OUT)
    out_ib, _ = O.pool_forward(A2_X, 2, True)
    assert np.array_equal(out_ib, A2_OUT_IB)
    _, pc = O.pool_forward(A3_X, 2, False)
    assert np.array_equal(O.pool_backward(A3_DOUT, pc), A3_DX)
    lp = O.log_softmax(A4_Z)
    assert close(lp[0], A4_LOGP0, 1e-5)
    assert np.array_equal(np.argmax(O.log_softmax(A4_Z), axis=1), A4_PRED)
    assert close(O.act_forward('relu10', A5_Z), A5_A)
    assert close(O.act_backward('relu10', A5_Z, A5_A, np.ones(3, F32)), A5_SLOPE)
    assert np.array_equal(O._iround(A6_IN), A6_ROUND)
    reg = dict(O.DEFAULT_REG, momentum=A7['m'], rate=1, maxnorm=0)
    t2, v2 = O.sgd_update(np.array([A7['theta']], F32), np.array([A7['v']], F32), np.array([A7['g']], F32),
                          reg, F32(A7['lr']))
    assert close(t2, A7['theta_new']) and close(v2, A7['v_new'])


# --------------------------------------------------------------------------------------------
# the Theano stand-in the reference's own code runs over (tests/golden/make_golden_ref.py)
# --------------------------------------------------------------------------------------------
@pytest.fixture()
def shim():
    path = os.path.join(ROOT, 'oracle', 'theano_shim')
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'theano' or k.startswith('theano.')}
    sys.path.insert(0, path)
    import theano
    import theano.tensor as tt
    from theano.tensor.signal.pool import pool_2d
    from theano.tensor.nnet import conv2d
    yield theano, tt, pool_2d, conv2d
    sys.path.remove(path)
    for k in [k for k in sys.modules if k == 'theano' or k.startswith('theano.')]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_shim_against_hand_derived_vectors(shim):
    theano, tt, pool_2d, conv2d = shim
    # gradients are taken with respect to shared variables, as the reference does (layer.py:83)
    x, w = tt.tensor4('x'), theano.shared(A1_W)
    y = conv2d(x, w, border_mode='valid')
    out, dW = theano.function([x], [y, tt.grad(tt.sum(y * tt.constant(A1_G)), w)])(A1_X)
    assert np.array_equal(np.asarray(out), A1_OUT) and np.array_equal(np.asarray(dW), A1_DW)
    assert np.array_equal(np.asarray(theano.function([x], pool_2d(x, (2, 2), ignore_border=False))(A2_X)), A2_OUT)
    assert np.array_equal(np.asarray(theano.function([x], pool_2d(x, (2, 2), ignore_border=True))(A2_X)),
                          A2_OUT_IB)
    xs = theano.shared(A3_X)
    gx = theano.function([], tt.grad(tt.sum(pool_2d(xs, (2, 2), ignore_border=False) * tt.constant(A3_DOUT)),
                                     xs))()
    assert np.array_equal(np.asarray(gx), A3_DX)
    m = tt.matrix('m')
    lp = theano.function([m], tt.log(tt.nnet.softmax(m)))(A4_Z)
    assert close(np.asarray(lp)[0], A4_LOGP0, 1e-5)
    assert np.array_equal(np.asarray(theano.function([m], tt.argmax(m, axis=1))(A4_Z)), A4_PRED)
    vs = theano.shared(A5_Z)
    relu10 = tt.maximum(0, vs) + tt.minimum(0, vs) * 10 / 100            # layer.py:36 with NN = 10
    a, slope = theano.function([], [relu10, tt.grad(tt.sum(relu10), vs)])()
    assert close(np.asarray(a), A5_A) and close(np.asarray(slope), A5_SLOPE)
    v = tt.vector('v')
    r = theano.function([v], tt.iround(v))(A6_IN.astype(np.float64))
    assert np.array_equal(np.asarray(r), A6_ROUND)
    th = theano.shared(np.array([A7['theta']], F32))
    vel = theano.shared(np.array([A7['v']], F32))
    g = tt.vector('g')
    step = theano.function([g], [], updates=[(vel, A7['m'] * vel + (1 - A7['m']) * g),
                                             (th, th - A7['lr'] * vel)])
    step(np.array([A7['g']], F32))
    assert close(th.get_value(), A7['theta_new']) and close(vel.get_value(), A7['v_new'])


# --------------------------------------------------------------------------------------------
# the CUDA kernels, through the C ABI
# --------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_kernels_against_hand_derived_vectors():
    import torch
    from theanet_b200 import _C as C

    keep = []            # C.ptr() hands out raw pointers: the tensors must outlive the launches

    def dev(a, dtype=None):
        t = torch.from_numpy(np.ascontiguousarray(a))
        t = (t.to(dtype) if dtype is not None else t).cuda()
        keep.append(t)
        return t

    lin = C.act_code('linear')
    # A1: forward (generic direct kernel and the fused conv+pool kernels' conv stage) and dW
    xd, Wd, bd = dev(A1_X), dev(A1_W), dev(np.zeros(1, F32))
    out = torch.zeros((1, 1, 2, 2), device='cuda')
    C.call('tn_conv2d_fprop', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out), 1, 1, 4, 1, 3, 0, 2, *lin, None)
    assert np.array_equal(out.cpu().numpy(), A1_OUT)
    dW, db = torch.zeros_like(Wd), torch.zeros_like(bd)
    ws = torch.zeros(4096, device='cuda')
    C.call('tn_conv2d_wgrad', C.ptr(xd), C.ptr(dev(A1_G)), C.ptr(dW), C.ptr(db), C.ptr(ws), 1, 1, 4, 1, 3,
           0, 2, None)
    assert np.array_equal(dW.cpu().numpy(), A1_DW)
    # ... and through the fused conv + 2x2 pool training kernel: pooled = max(10, 13, 22, 25) = 25,
    # un-pooled activations = A1_OUT
    a = torch.zeros((1, 1, 2, 2), device='cuda')
    pooled = torch.zeros((1, 1, 1, 1), device='cuda')
    tie = torch.zeros(1, dtype=torch.uint8, device='cuda')
    if C.lib.tn_convpool_small_supported(1, 4, 1, 3, 0, 2, lin[0], 2, 1):
        C.call('tn_convpool_fprop_train', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(a), C.ptr(pooled),
               C.ptr(tie), 1, 1, 4, 1, 3, 0, 2, *lin, 2, 1, None)
        assert np.array_equal(a.cpu().numpy(), A1_OUT) and pooled.item() == 25 and tie.item() == 8
    # A2 / A3
    for ib, want in ((False, A2_OUT), (True, A2_OUT_IB)):
        o = want.shape[-1]
        po = torch.zeros((1, 1, o, o), device='cuda')
        C.call('tn_maxpool_fwd', C.ptr(dev(A2_X)), C.ptr(po), 1, 5, 2, o, None)
        assert np.array_equal(po.cpu().numpy(), want)
    x3 = dev(A3_X)
    p3 = torch.zeros((1, 1, 2, 2), device='cuda')
    C.call('tn_maxpool_fwd', C.ptr(x3), C.ptr(p3), 1, 3, 2, 2, None)
    dx3 = torch.zeros((1, 1, 3, 3), device='cuda')
    C.call('tn_maxpool_bwd', C.ptr(dev(A3_DOUT)), C.ptr(x3), C.ptr(p3), C.ptr(dx3), 1, 3, 2, 2, *lin, None)
    assert np.array_equal(dx3.cpu().numpy(), A3_DX)
    # A4
    ctl = dev(np.zeros(C.CTL_WORDS, np.int32))
    y = dev(np.array([0, 0], np.int32))
    lp = torch.zeros((2, 3), device='cuda')
    preds = torch.zeros(2, dtype=torch.int64, device='cuda')
    stats = torch.zeros(2 + 2 * 2, device='cuda')
    C.call('tn_softmax_test_stats', C.ptr(dev(A4_Z)), C.ptr(y), None, C.ptr(ctl), 2, 3, C.ptr(lp),
           C.ptr(preds), C.ptr(stats), None)
    assert close(lp.cpu().numpy()[0], A4_LOGP0, 1e-5)
    assert np.array_equal(preds.cpu().numpy(), A4_PRED)
    # A5: forward through the dense kernel's epilogue (x = I, W = z as a 1x3... use act_bwd for the slope)
    gz = torch.zeros(3, device='cuda')
    C.call('tn_act_bwd', C.ptr(dev(np.ones(3, F32))), C.ptr(dev(A5_A)), C.ptr(gz), 3, *C.act_code('relu10'), None)
    torch.cuda.synchronize()
    assert close(gz.cpu().numpy(), A5_SLOPE), gz.cpu().numpy()
    # A7
    seg = (C.ParamSeg * 1)()
    s = seg[0]
    s.offset, s.size, s.ndim, s.rows, s.cols = 0, 1, 1, 1, 1
    s.momentum, s.rate, s.maxnorm, s.l1, s.l2 = A7['m'], 1.0, 0.0, 0.0, 0.0
    th, vel, g = dev(np.array([A7['theta'], 0, 0, 0], F32)), dev(np.array([A7['v'], 0, 0, 0], F32)), \
        dev(np.array([A7['g'], 0, 0, 0, 0, 0, 0, 0], F32))
    c = np.zeros(C.CTL_WORDS, np.int32)
    c[C.CTL_LR_BITS] = F32(A7['lr']).view(np.int32)
    nb = C.lib.tn_update_workspace_bytes(1, 4)
    wsu = torch.zeros(nb // 4 + 1, device='cuda')
    cost = torch.zeros(1, device='cuda')
    C.call('tn_sgd_momentum_maxnorm_update', C.ptr(th), C.ptr(vel), C.ptr(g), seg, 1, 4, C.ptr(dev(c)), 1.0,
           None, 1.0, C.ptr(cost), C.ptr(wsu), None)
    torch.cuda.synchronize()
    assert close(th.cpu().numpy()[0], A7['theta_new']) and close(vel.cpu().numpy()[0], A7['v_new'])
