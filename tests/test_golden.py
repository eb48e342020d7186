"""Golden vectors produced by tools/make_golden.py from the oracle with the product's Philox streams
(the vectors recorded from the reference's own code are in tests/test_golden_ref.py).  CPU: the oracle
still reproduces them.  GPU (-m gpu): the CUDA path reproduces them through the public API."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from oracle import theanet_oracle as O   # noqa: E402
import make_golden as MG                 # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
CASES = [('mnist_b8', 'mnist.prms', 10, 4), ('flat3_b8', '3flat.prms', 457, 3)]
TOL = 1e-3      # BASELINE.json north_star: 1e-3 relative in float32


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def check_digest(t, want, tol):
    got = MG.digest(t)
    assert abs(got[0] - want[0]) <= tol * max(abs(want[0]), np.sqrt(want[1]))      # sum
    assert abs(got[1] - want[1]) <= 2 * tol * want[1] + 1e-30                       # sum of squares
    assert rel(got[3:], want[3:]) < tol or want[2] < 1e-12


@pytest.mark.parametrize('name,prms_file,ncls,steps', CASES)
def test_oracle_reproduces_golden(name, prms_file, ncls, steps):
    g = np.load(os.path.join(GOLD, name + '.npz'))
    B, img = int(g['B']), int(g['img'])
    p = MG.load_prms(prms_file, B, img)
    on = O.OracleNet(p['layers'], p['training_params'])
    for s in range(steps):
        i = s % 2
        cost, lp = on.train_step(g['x'][i * B:(i + 1) * B], g['y'][i * B:(i + 1) * B], step=s, sample0=0)
        assert abs(cost - g['cost_%d' % s]) <= 1e-6 * abs(g['cost_%d' % s])
        assert rel(lp, g['logprob_%d' % s]) < 1e-6
    k = 0
    for L in on.spec:
        for j, t in enumerate(L['params'] or []):
            check_digest(t, g['w_%d' % k], 1e-6)
            check_digest(L['vel'][j], g['v_%d' % k], 1e-6)
            k += 1


def test_oracle_kernel_vectors():
    g = np.load(os.path.join(GOLD, 'kernels.npz'))
    po, cache = O.pool_forward(g['pool_x'], 2, False)
    assert np.array_equal(po, g['pool_out'])
    assert np.array_equal(O.pool_backward(g['pool_dout'], cache), g['pool_dx'])
    z, cc = O.conv_forward(g['conv_x'], g['conv_w'], 'same')
    assert rel(z, g['conv_z']) < 1e-6
    dW, db, dx = O.conv_backward(g['conv_g'], g['conv_w'], cc)
    assert rel(dW, g['conv_dW']) < 1e-6 and rel(db, g['conv_db']) < 1e-6 and rel(dx, g['conv_dx']) < 1e-6
    reg = {"L1": 0, "L2": 0, "momentum": .9, "rate": 1, "maxnorm": 1.}
    th2, v2 = O.sgd_update(g['up_th'], g['up_vel'], g['up_gr'], reg, np.float32(.1))
    assert np.array_equal(th2, g['up_th2']) and np.array_equal(v2, g['up_vel2'])
    assert np.all(th2[:, 2] == 0)          # zero-norm column: scale (1e-7)/(1e-7) = 1


@pytest.mark.gpu
@pytest.mark.parametrize('name,prms_file,ncls,steps', CASES)
def test_gpu_reproduces_golden(name, prms_file, ncls, steps):
    from theanet_b200.neuralnet import NeuralNet
    g = np.load(os.path.join(GOLD, name + '.npz'))
    B, img = int(g['B']), int(g['img'])
    p = MG.load_prms(prms_file, B, img)
    net = NeuralNet(p['layers'], p['training_params'])
    fn = net.get_trin_model(g['x'], g['y'])
    for s in range(steps):
        cost, feats, lp = fn(s % 2)
        assert abs(cost - g['cost_%d' % s]) <= TOL * abs(g['cost_%d' % s])
        assert rel(lp, g['logprob_%d' % s]) < TOL
    k = 0
    vel = net.get_velocities()
    for li, ww in enumerate(net.get_init_params()['allwts']):
        for j, t in enumerate(ww):
            check_digest(t, g['w_%d' % k], TOL)
            check_digest(vel[li][j], g['v_%d' % k], TOL)
            k += 1
    e, pr, lp, yp = net.get_test_model(g['x'], g['y'], preds_feats=True)(0)
    assert abs(e - g['test_err']) < 1e-6 and abs(pr - g['test_py']) <= TOL * g['test_py']
    assert np.array_equal(yp, g['test_pred'])              # argmax: bit-exact index work
