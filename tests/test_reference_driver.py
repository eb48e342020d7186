"""The reference's OWN driver on this engine.

INTEGRATION.md claims that /root/reference/train.py runs against theanet_b200 after three edits.
Here the edits are applied programmatically to the reference's file as it lies (read at test time,
never copied into the repository):

  * CPU: every hunk matches exactly once, the result compiles, and everything the driver touches
    on `nn` / `net` exists with the reference's signature; a `.pkl` written from get_init_params()
    is digested by the reference's print_pkl_info.py.
  * GPU (-m gpu): the patched driver trains one epoch of params/mnist.prms on data/synthetic, tests,
    saves its `.pkl`, and print_pkl_info.py reads that file.

All of it is skipped where /root/reference does not exist (the GPU box)."""
import ast
import inspect
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'train.py')),
                               reason="the reference tree is not present on this machine")

# INTEGRATION.md section 1, verbatim
HUNKS = [
    ("import theano as th\nimport theanet.neuralnet as nn\n",
     "import theanet_b200.neuralnet as nn\n"),
    ("def share(data, dtype=th.config.floatX, borrow=True):\n"
     "    return th.shared(np.asarray(data, dtype), borrow=borrow)\n",
     "def share(data, dtype='float32', borrow=True):\n"
     "    return np.asarray(data, dtype)\n"),
    ("print('Device : {} ({})'.format(th.config.device, th.config.floatX))\n",
     "print('Device : cuda (float32)')\n"),
]


def patched_driver():
    with open(os.path.join(REF, 'train.py')) as f:
        src = f.read()
    for old, new in HUNKS:
        assert src.count(old) == 1, "hunk does not apply exactly once:\n" + old
        src = src.replace(old, new)
    assert 'theano' not in src and 'th.' not in src.replace('path.', '').replace('with ', '')
    return src


@needs_ref
def test_the_three_edits_apply_and_the_api_surface_exists():
    src = patched_driver()
    tree = ast.parse(src)                                   # the patched driver is valid Python
    import theanet_b200.neuralnet as nn
    # module-level names the driver uses on `nn`
    used_nn = {n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and
               isinstance(n.value, ast.Name) and n.value.id == 'nn'}
    assert used_nn == {'NeuralNet', 'get_layers_info', 'get_training_params_info'}
    for name in used_nn:
        assert hasattr(nn, name), name
    # methods called on `net`
    used_net = {n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and
                isinstance(n.value, ast.Name) and n.value.id == 'net'}
    assert used_net == {'get_trin_model', 'get_test_model', 'tr_layers', 'get_init_params', 'get_wts_info',
                        'get_epoch', 'inc_epoch_set_rate'}
    for name in used_net - {'tr_layers'}:
        assert callable(getattr(nn.NeuralNet, name)), name
    # call shapes: NeuralNet(layers, tr_prms, allwts); get_trin_model(x, y, aux); get_test_model(x, y, aux)
    assert list(inspect.signature(nn.NeuralNet.__init__).parameters)[1:4] == ['layers', 'training_params', 'allwts']
    assert list(inspect.signature(nn.NeuralNet.get_trin_model).parameters)[1:4] == ['x_data', 'y_data', 'aux_data']
    assert list(inspect.signature(nn.NeuralNet.get_test_model).parameters)[1:4] == ['x_data', 'y_data', 'aux_data']
    assert 'detailed' in inspect.signature(nn.NeuralNet.get_wts_info).parameters


def small_prms(path, batch=64, seed=4711):
    with open(os.path.join(ROOT, 'params', 'mnist.prms')) as f:
        p = ast.literal_eval(f.read())
    p['training_params'].update(SEED=seed, BATCH_SZ=batch, NUM_EPOCHS=1, TEST_SAMP_SZ=4 * batch)
    with open(path, 'w') as f:
        f.write(repr(p))
    return p


# print_pkl_info.py indexes a str with the numpy.bool of `np.prod(...) == 1`, which NumPy >= 2 refuses
# (the script predates it).  It is executed UNMODIFIED with np.prod returning Python ints for scalars.
PKL_INFO_RUNNER = """
import runpy, sys
import numpy as np
_prod = np.prod
np.prod = lambda *a, **k: (lambda v: int(v) if np.ndim(v) == 0 else v)(_prod(*a, **k))
sys.argv = [sys.argv[1], sys.argv[2]]
runpy.run_path(sys.argv[0], run_name='__main__')
"""


def run_print_pkl_info(pkl, cwd):
    r = subprocess.run([sys.executable, '-c', PKL_INFO_RUNNER, os.path.join(REF, 'print_pkl_info.py'), pkl],
                       capture_output=True, text=True, cwd=cwd, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@needs_ref
def test_the_reference_pkl_inspector_reads_our_pkl(tmp_path):
    """get_init_params() on a host-side net (no kernels run) -> pickle -> print_pkl_info.py."""
    from theanet_b200.neuralnet import NeuralNet
    p = small_prms(str(tmp_path / 'm.prms'))
    p['layers'][0][1]['img_sz'] = 28
    net = NeuralNet(p['layers'], p['training_params'], device='cpu')
    pkl = str(tmp_path / 'net.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(net.get_init_params(), f, -1)
    out = run_print_pkl_info(pkl, str(tmp_path))
    n_wts = sum(int(np.prod(w.shape)) for wb in net.get_init_params()['allwts'] for w in wb)
    assert "Total Number of Weights: {:,}".format(n_wts) in out
    for name in ('ElasticLayer', 'ConvLayer', 'PoolLayer', 'HiddenLayer', 'SoftmaxLayer'):
        assert name in out
    assert "Shape:(4, 1, 3, 3) = 36" in out and "Shape:(720, 500) = 360,000" in out


@needs_ref
@pytest.mark.gpu
def test_the_reference_driver_trains_an_epoch_on_this_engine(tmp_path):
    drv = str(tmp_path / 'train_reference_patched.py')
    with open(drv, 'w') as f:
        f.write(patched_driver())
    small_prms(str(tmp_path / 'mnistsmall.prms'))
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''),
               TN_SYNTH_TRAIN='2048', TN_SYNTH_TEST='512')
    r = subprocess.run([sys.executable, '-W', 'ignore', drv, 'synthetic', 'mnistsmall.prms'], capture_output=True,
                       text=True, cwd=str(tmp_path), env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = r.stdout
    assert 'Device : cuda (float32)' in out and 'Training ...' in out
    assert 'Final Error Rates' in out or 'Te_Error' in out
    epoch_lines = [l for l in out.splitlines() if l.strip().startswith('0 ') and '%' in l]
    assert epoch_lines, out[-2000:]                          # "  0   <cost>    tr%  (p)   te%  (p)"
    cost = float(epoch_lines[0].split()[1])
    assert np.isfinite(cost) and cost > 0
    pkls = [f for f in os.listdir(str(tmp_path)) if f.endswith('.pkl')]
    assert pkls, os.listdir(str(tmp_path))
    info = run_print_pkl_info(os.path.join(str(tmp_path), pkls[0]), str(tmp_path))
    assert "Total Number of Weights: 365,810" in info or "Total Number of Weights" in info
