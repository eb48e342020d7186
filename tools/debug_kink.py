"""Scratch diagnosis (GPU): does the C5 first step differ from the oracle only through ReLU-kink flips?"""
import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import theanet_oracle as O
from theanet_b200 import _C
from theanet_b200.neuralnet import NeuralNet
import test_gpu_net as T
prms = T.load_prms('mnist.prms', 512, 64)
x, y = T.synth(1024, 1, 64, 10)
for mode in (0, 1):
    _C.call('tn_set_dense_mode', mode)
    p1, p2 = copy.deepcopy(prms), copy.deepcopy(prms)
    net = NeuralNet(p1['layers'], p1['training_params'], use_graph=False)
    on = O.OracleNet(p2['layers'], p2['training_params'])
    fn = net.get_trin_model(x, y)
    fn(0)
    on.train_step(x[:512], y[:512], step=0, sample0=0, apply_update=False)
    for li in (1, 3, 5):
        a_dev = net.out[li].cpu().numpy().reshape(512, -1) if net.out[li] is not None else None
        c = on.last_caches[li]
        if a_dev is None or 'z' not in c:
            print('layer', li, 'no data', None if a_dev is None else a_dev.shape, list(c.keys())); continue
        z = c['z'].reshape(512, -1); a_or = c['out'].reshape(512, -1)
        m = c.get('mask', np.ones_like(z)).reshape(512, -1)
        flip = ((a_dev > 0) != (a_or > 0)) & (m != 0)
        print('mode', mode, 'layer', li, 'flips', int(flip.sum()), 'of', flip.size, 'z at flips', z[flip][:8],
              'dev a', a_dev[flip][:8], 'max|z|', np.abs(z).max(), 'n |z|<1e-6', int((np.abs(z) < 1e-6).sum()))
    g_dev = net.dbuf[5].cpu().numpy()
    print('   done mode', mode)
