"""Generate the golden vectors under tests/golden/ from the CPU oracle.

These fixtures use the PRODUCT's Philox random streams, which only the oracle (not the reference)
can reproduce; the vectors recorded from the reference's own code, with its own draws, are made by
tests/golden/make_golden_ref.py.  They freeze the oracle against regressions (tests/test_golden.py,
CPU) and give the GPU path a fixed target for networks with random layers (-m gpu).

    python tools/make_golden.py            # rewrites tests/golden/*.npz
"""
import ast
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import theanet_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def load_prms(name, B, img, seed=555555):
    with open(os.path.join(ROOT, 'params', name)) as f:
        p = ast.literal_eval(f.read())
    p['training_params'].update(SEED=seed, BATCH_SZ=B)
    p['layers'][0][1]['img_sz'] = img
    return p


def synth(n, c, s, ncls, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, c, s, s), dtype=np.float32)
    x *= (x > .8)                                     # ~80% exact zeros: pool ties
    y = rng.integers(0, ncls, n).astype(np.int32)
    return x, y


def digest(t, n=256):
    """Small fingerprint of a tensor: [sum, sum of squares, max |.|] in float64 followed by `n`
    elements at fixed pseudo-random positions (fixtures stay a few kB per tensor)."""
    f = np.asarray(t, np.float64).ravel()
    idx = np.random.default_rng(f.size).integers(0, f.size, n)
    return np.concatenate([[f.sum(), (f * f).sum(), np.abs(f).max()], f[idx]])


def whole_step(name, prms_file, B, img, ncls, steps):
    p = load_prms(prms_file, B, img)
    x, y = synth(2 * B, 1, img, ncls, 1234)
    on = O.OracleNet(copy.deepcopy(p['layers']), copy.deepcopy(p['training_params']))
    out = {'x': x, 'y': y, 'B': np.int32(B), 'img': np.int32(img)}
    for s in range(steps):
        i = s % 2
        cost, lp = on.train_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B], step=s, sample0=0)
        out['cost_%d' % s] = np.float32(cost)
        out['logprob_%d' % s] = lp.astype(np.float32)
    k = 0
    for L in on.spec:
        for j, t in enumerate(L['params'] or []):
            out['w_%d' % k] = digest(t)
            out['v_%d' % k] = digest(L['vel'][j])
            k += 1
    e, pr, lp, yp = on.test_step(x[:B], y[:B])
    out['test_err'], out['test_py'], out['test_pred'] = np.float32(e), np.float32(pr), yp.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'cost', [float(out['cost_%d' % s]) for s in range(steps)])


def kernel_vectors():
    rng = np.random.default_rng(99)
    out = {}
    # pool: ties inside windows, odd size (1-wide edge windows), all-negative windows
    x = (rng.integers(-3, 2, (2, 3, 11, 11)) / 2).astype(np.float32)
    x[0, 0, :4, :4] = -1.5
    po, cache = O.pool_forward(x, 2, False)
    dout = rng.standard_normal(po.shape).astype(np.float32)
    out.update(pool_x=x, pool_out=po, pool_dout=dout, pool_dx=O.pool_backward(dout, cache))
    # conv 'same' + leaky relu
    cx = rng.standard_normal((2, 3, 9, 9)).astype(np.float32)
    cw = (rng.standard_normal((5, 3, 3, 3)) / 5).astype(np.float32)
    cz, cc = O.conv_forward(cx, cw, 'same')
    cg = rng.standard_normal(cz.shape).astype(np.float32)
    dW, db, dx = O.conv_backward(cg, cw, cc)
    out.update(conv_x=cx, conv_w=cw, conv_z=cz, conv_g=cg, conv_dW=dW, conv_db=db, conv_dx=dx)
    # elastic: nearest with translation/zoom/rotation + smoothing
    prm = dict(translation=2, zoom=1.1, magnitude=30, sigma=4, pflip=0., angle=15,
               invert_image=True, nearest=True)
    h = 16
    noise = rng.standard_normal((2, h, h)).astype(np.float32)
    u = rng.random(8).astype(np.float32)
    ty, tx = O.elastic_target(h, prm, noise, u)[-2:]
    ex = rng.random((2, 1, h, h)).astype(np.float32)
    out.update(el_noise=noise, el_u=u, el_ty=ty, el_tx=tx, el_x=ex,
               el_out=O.elastic_apply(ex, prm, ty, tx))
    # update: maxnorm on columns with a zero-norm column, lagged momentum
    th = rng.standard_normal((6, 5)).astype(np.float32)
    th[:, 2] = 0
    vel = rng.standard_normal((6, 5)).astype(np.float32) * .1
    vel[:, 2] = 0
    gr = rng.standard_normal((6, 5)).astype(np.float32)
    gr[:, 2] = 0
    reg = {"L1": 0, "L2": 0, "momentum": .9, "rate": 1, "maxnorm": 1.}
    th2, vel2 = O.sgd_update(th, vel, gr, reg, np.float32(.1))
    out.update(up_th=th, up_vel=vel, up_gr=gr, up_th2=th2, up_vel2=vel2)
    np.savez_compressed(os.path.join(OUT, 'kernels.npz'), **out)
    print('kernels.npz', sorted(out))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    whole_step('mnist_b8', 'mnist.prms', 8, 28, 10, 4)
    whole_step('flat3_b8', '3flat.prms', 8, 28, 457, 3)
    kernel_vectors()
