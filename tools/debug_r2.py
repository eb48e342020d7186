"""Scratch diagnosis (GPU): per-tensor deviations of one training step from the oracle under
different kernel selections.   python tools/debug_r2.py [c5] [b8]"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from oracle import theanet_oracle as O          # noqa: E402
from theanet_b200 import _C                     # noqa: E402
from theanet_b200.neuralnet import NeuralNet    # noqa: E402
import test_gpu_net as T                        # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def one(prms, x, y, steps=1, tag='', **kw):
    p1, p2 = copy.deepcopy(prms), copy.deepcopy(prms)
    net = NeuralNet(p1['layers'], p1['training_params'], **kw)
    on = O.OracleNet(p2['layers'], p2['training_params'])
    B = prms['training_params']['BATCH_SZ']
    fn = net.get_trin_model(x, y)
    for s in range(steps):
        cost, _, lp = fn(s)
        ocost, olp = on.train_step(x[s * B:(s + 1) * B], y[s * B:(s + 1) * B], step=s, sample0=0)
    out = []
    for li, (vs, L) in enumerate(zip(net.get_velocities(), on.spec)):
        for k, u in enumerate(vs):
            out.append('L{}.{}:{:.1e}'.format(li, k, rel(u, L['vel'][k])))
    print(tag, 'cost', float(cost), float(ocost), ' '.join(out), flush=True)


def c5():
    prms = T.load_prms('mnist.prms', 512, 64)
    x, y = T.synth(1024, 1, 64, 10)
    one(prms, x, y, tag='c5 default      ')
    one(prms, x, y, tag='c5 no graph     ', use_graph=False)
    os.environ['TN_OVERLAP_WGRAD'] = '0'
    one(prms, x, y, tag='c5 no overlap   ')
    os.environ['TN_OVERLAP_WGRAD'] = '1'
    for mode in (1, 2):
        _C.call('tn_set_dense_mode', mode)
        one(prms, x, y, tag='c5 dense mode {}  '.format(mode))
    _C.call('tn_set_dense_mode', 0)
    one(prms, x, y, tag='c5 no fuse_conv ', fuse_conv=False)
    for sp in ('2', '4'):
        os.environ['TN_SK_SPLIT'] = sp
        one(prms, x, y, tag='c5 split {}      '.format(sp))
    os.environ.pop('TN_SK_SPLIT')
    os.environ['TN_SK_BN'] = '64'
    one(prms, x, y, tag='c5 BN 64        ')
    os.environ.pop('TN_SK_BN')


def b8():
    import make_golden as MG
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'mnist_b8.npz'))
    B, img = int(g['B']), int(g['img'])
    p = MG.load_prms('mnist.prms', B, img)
    for kw in ({}, {'fuse_conv': False}, {'use_graph': False}):
        for steps in (1, 2, 4):
            one(p, g['x'], g['y'], steps=min(steps, 2), tag='b8 {} steps {}'.format(kw, min(steps, 2)), **kw)
    # tie localisation on the golden corpus
    net, nd, nw = T.run_tie_localised(p, g['x'], g['y'], 2, pools=(2, 4))
    print('b8 tie-localised ok; windows differing', nd, 'of', nw)


if __name__ == '__main__':
    which = sys.argv[1:] or ['c5', 'b8']
    if 'c5' in which:
        c5()
    if 'b8' in which:
        b8()
