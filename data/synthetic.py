"""MNIST-shaped synthetic corpus (no network in this environment; the reference's data/mnist.py
downloads its pickle).  Same module protocol as the reference's data modules: training_x,
training_y, testing_x, testing_y.  Classes are separable blobs so that training visibly learns."""
import os

import numpy as np

N_TRAIN = int(os.environ.get('TN_SYNTH_TRAIN', 8192))
N_TEST = int(os.environ.get('TN_SYNTH_TEST', 2048))
IMG = int(os.environ.get('TN_SYNTH_IMG', 28))
N_CLASSES = 10


def _make(n, rng, protos):
    y = rng.integers(0, N_CLASSES, n).astype(np.int32)
    x = protos[y] + .35 * rng.standard_normal((n, 1, IMG, IMG)).astype(np.float32)
    x = np.clip(x, 0, 1)
    x *= (x > .45)                 # mostly exact zeros, as MNIST
    return x.astype(np.float32), y


_rng = np.random.default_rng(1234)
_protos = (_rng.random((N_CLASSES, 1, IMG, IMG)) > .8).astype(np.float32)
training_x, training_y = _make(N_TRAIN, _rng, _protos)
testing_x, testing_y = _make(N_TEST, _rng, _protos)
