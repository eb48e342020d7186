"""The only real image data available offline: sklearn's 8x8 digits, zero-padded to 28x28 so that
params/mnist.prms runs on it unchanged.  Same protocol as the reference's data modules."""
import numpy as np
from sklearn.datasets import load_digits

_d = load_digits()
_x = (_d.images / 16.).astype(np.float32)
_x = np.kron(_x, np.ones((3, 3), np.float32))                 # 24x24
_x = np.pad(_x, ((0, 0), (2, 2), (2, 2)))[:, None, :, :]       # (N, 1, 28, 28)
_y = _d.target.astype(np.int32)
_perm = np.random.default_rng(0).permutation(len(_x))
_x, _y = _x[_perm], _y[_perm]
_n = 1500
training_x, training_y = _x[:_n], _y[:_n]
testing_x, testing_y = _x[_n:], _y[_n:]
