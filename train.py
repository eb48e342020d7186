#!/usr/bin/env python
"""Training driver with the command line and console output of the reference's train.py
(reference: train.py:59-245), running on theanet_b200.

    python train.py <dataset> <params_file> [redirect=0]

<dataset> is a module under data/ exposing training_x/_y and testing_x/_y (data/synthetic.py,
data/digits.py; the reference's data/mnist.py needs the network).  <params_file> is a .prms
(python literal) or a .pkl written by a previous run.  What differs from the reference driver:
no `import theano` (the corpus is handed over as plain arrays and becomes HBM-resident; the
reference wraps it in theano.shared, train.py:18-19,126-129), and `sys.argv[-1] == '1'` instead of
the reference's identity comparison (train.py:100).
"""
import ast
import importlib
import os
import pickle
import socket
import sys
from datetime import datetime

import numpy as np


def load_params(path):
    if path.endswith('.pkl'):
        with open(path, 'rb') as f:
            return pickle.load(f)
    with open(path) as f:
        return ast.literal_eval(f.read())


def as_images(arr):
    """(N, S*S) / (N, S, S) / (N, C, S, S) -> (N, C, S, S), cf. train.py:22-34."""
    arr = np.asarray(arr)
    if arr.ndim == 2:
        side = int(round(arr.shape[1] ** .5))
        if side * side != arr.shape[1]:
            raise ValueError("flat images need a perfect-square length")
        return arr.reshape(arr.shape[0], 1, side, side)
    if arr.ndim == 3:
        return arr[:, None]
    if arr.ndim == 4:
        return arr
    raise ValueError("image arrays must have 2, 3 or 4 dimensions")


class Tee:
    """stdout, optionally redirected to <params>_<SEED>.txt (train.py:37-55,100-104)."""

    def __init__(self, path=None):
        self.path = path
        self.stream = open(path, 'w', 1) if path else sys.__stdout__

    def write(self, s):
        self.stream.write(s)

    def flush(self):
        self.stream.flush()

    def forceflush(self):
        if self.path:
            self.stream.close()
            self.stream = open(self.path, 'a', 1)


def window_indices(total, batch_sz, window):
    """Rotating window of batch indices used for the periodic tests (train.py:170-176)."""
    each, n_all, cur = int(window / batch_sz), int(total / batch_sz), 0
    while True:
        yield [i % n_all for i in range(cur, cur + each)]
        cur = (cur + each) % n_all


def epoch_batches(n, batch_sz, rng=None):
    """What one epoch feeds the training function: batch indices 0..n//B-1 in order (train.py:210),
    or, with a RandomState, n//B index vectors of a fresh permutation (remainder dropped)."""
    nb = n // batch_sz
    if rng is None:
        return list(range(nb))
    perm = rng.permutation(n)[:nb * batch_sz].astype(np.int32)
    return [perm[i * batch_sz:(i + 1) * batch_sz] for i in range(nb)]


def percent_errors(pairs):
    pairs = list(pairs)
    return tuple(100 * float(np.mean([p[k] for p in pairs])) for k in (0, 1))


def main(argv):
    if len(argv) < 3:
        print("Usage: {} <dataset> <params_file(.prms|.pkl)> [redirect=0]".format(argv[0]))
        return 1
    import theanet_b200.neuralnet as nn

    dataset_name, prms_path = argv[1], argv[2]
    params = load_params(prms_path)
    layers, tr_prms = params['layers'], params['training_params']
    allwts = params.get('allwts')
    if tr_prms.get('SEED') is None:
        tr_prms['SEED'] = int(np.random.randint(0, 10 ** 6))
    head = os.path.splitext(os.path.basename(prms_path))[0] + "_{:06d}".format(tr_prms['SEED'])
    out = Tee(head + '.txt' if argv[-1] == '1' else None)
    sys.stdout = out

    print(' '.join(argv))
    print('Time   :' + datetime.now().strftime('%Y-%m-%d %H:%M:%S'))
    print('Device : cuda (float32) theanet_b200')
    print('Host   :', socket.gethostname())
    print(nn.get_layers_info(layers))
    print(nn.get_training_params_info(tr_prms))

    data = importlib.import_module('data.' + dataset_name)
    trin_x, test_x = as_images(data.training_x), as_images(data.testing_x)
    n_tr, _, _, layers[0][1]['img_sz'] = trin_x.shape
    n_te = test_x.shape[0]

    print("\nInitializing the net ... ")
    net = nn.NeuralNet(layers, tr_prms, allwts)
    if params.get('resume') is not None:          # written by this driver: momentum, step, seeds
        net.set_resume_state(params['resume'])
    print(net)
    print(net.get_wts_info(detailed=True).replace("\n\t", ""))

    print("\nCompiling ... ")
    # 'SHUFFLE': True (an extension; the reference's TODO:14-16 asks for it) draws a fresh
    # permutation of the training set every epoch and feeds index lists (neuralnet.py:228-234);
    # the default walks the batches in their fixed order like the reference (train.py:210)
    shuffle = bool(tr_prms.get('SHUFFLE', False))
    # auxiliary inputs when the net wants them and the dataset has them (reference train.py:131-135)
    trin_aux = getattr(data, 'training_aux', None) if net.takes_aux() else None
    test_aux = getattr(data, 'testing_aux', None) if net.takes_aux() else None
    training_fn = net.get_trin_model(trin_x, data.training_y, trin_aux, take_index_list=shuffle)
    order_rng = np.random.RandomState(tr_prms['SEED'] + 1)
    test_fn_tr = net.get_test_model(trin_x, data.training_y, trin_aux)
    test_fn_te = net.get_test_model(test_x, data.testing_y, test_aux)

    B, n_epochs = tr_prms['BATCH_SZ'], tr_prms['NUM_EPOCHS']
    aux_name = 'BitErr' if net.tr_layers[-1].kind == 'LOGIT' else 'P(MLE)'
    te_idx = window_indices(n_te, B, tr_prms['TEST_SAMP_SZ'])
    tr_idx = window_indices(n_tr, B, tr_prms['TEST_SAMP_SZ'])
    saved = [None]

    def do_test():
        te = percent_errors(test_fn_te(i) for i in next(te_idx))
        tr = percent_errors(test_fn_tr(i) for i in next(tr_idx))
        print("{:5.2f}%  ({:5.2f}%)      {:5.2f}%  ({:5.2f}%)".format(tr[0], tr[1], te[0], te[1]))
        out.forceflush()
        if saved[0]:
            os.remove(saved[0])
        saved[0] = head + '_{:02.0f}.pkl'.format(te[0])
        with open(saved[0], 'wb') as f:
            # the reference's schema (neuralnet.py:298-301) + one extra key for an exact resume
            pickle.dump(dict(net.get_init_params(), resume=net.get_resume_state()), f, -1)

    print("Training ...")
    print("Epoch   Cost  Tr_Error Tr_{0}    Te_Error Te_{0}".format(aux_name))
    for epoch in range(n_epochs):
        total_cost = 0.
        batches = epoch_batches(n_tr, B, order_rng if shuffle else None)
        for ibatch, rows in enumerate(batches):             # remainder dropped
            cost, features, logprobs = training_fn(rows)
            total_cost += cost
            if np.isnan(total_cost):
                print(net.get_wts_info(detailed=True))
                raise ZeroDivisionError("Nan cost at Epoch:{} Iteration:{}".format(epoch, ibatch))
        if epoch % tr_prms['EPOCHS_TO_TEST'] == 0:
            print("{:3d} {:>8.2f}".format(net.get_epoch(), total_cost), end='    ')
            do_test()
        net.inc_epoch_set_rate()

    te = percent_errors(test_fn_te(i) for i in range(n_te // B))
    tr = percent_errors(test_fn_tr(i) for i in range(n_tr // B))
    print("{:3d} {:>8.2f}".format(net.get_epoch(), 0), end='    ')
    print("{:5.2f}%  ({:5.2f}%)      {:5.2f}%  ({:5.2f}%)".format(tr[0], tr[1], te[0], te[1]))
    return 0


if __name__ == '__main__':
    sys.exit(main(sys.argv))
