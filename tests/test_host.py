"""Host logic of the drop-in (no GPU): the constructor, shape inference, weight initialisation,
printable representations and error behaviour of theanet_b200.NeuralNet against what the
REFERENCE'S OWN constructor produced for the same .prms and SEED (tests/golden/ref_*.npz, recorded
by tests/golden/make_golden_ref.py), plus the facade's CPU-side contract.

Nothing here computes on a device: NeuralNet builds its layer twins and buffers on the CPU and
refuses to run there (there is no CPU execution path)."""
import copy
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden_ref as MR             # noqa: E402
from theanet_b200 import neuralnet as nn  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def build(name):
    c = MR.CASES[name]
    g = np.load(os.path.join(GOLD, 'ref_%s.npz' % name))
    layers, tp = copy.deepcopy(c['layers']), copy.deepcopy(c['tp'])
    return c, g, nn.NeuralNet(layers, tp, device='cpu')


@pytest.mark.parametrize('name', sorted(MR.CASES))
def test_constructor_prints_what_the_reference_prints(name):
    """str(net) (both twins, train.py:139), the layer / training-parameter listings after the
    constructor's in-place edits (train.py:115-116; neuralnet.py:108-109,132-136) and the detailed
    weight summary (train.py:140), character for character."""
    c, g, net = build(name)
    assert str(net).split('\nParams')[0] == str(g['repr'])
    assert net.get_layers_info() == str(g['layers_info'])
    assert net.get_training_params_info() == str(g['tp_info'])
    assert net.get_wts_info(detailed=True) == str(g['wts_info'])


@pytest.mark.parametrize('name', sorted(MR.CASES))
def test_shape_inference_and_initial_weights_match_the_reference(name):
    """n_out / num_maps / out_sz of every layer (neuralnet.py:113-201) and the initial parameters:
    one numpy RandomState walks through the constructors in the reference's order (weights.py:40-68,
    the randint(1e6) of every random stream included), so they are bit-identical."""
    c, g, net = build(name)
    got = [[getattr(l, 'n_out', -1), getattr(l, 'num_maps', -1) or -1, getattr(l, 'out_sz', -1) or -1]
           for l in net.tr_layers]
    assert np.array_equal(np.array(got), g['shapes'])
    k = 0
    for ww in net.get_init_params()['allwts']:
        for t in ww:
            assert t.dtype == np.float32
            assert np.array_equal(MR.thin(t), g['w0_%d' % k])
            assert np.allclose(MR.digest(t), g['w0d_%d' % k], rtol=1e-12)
            k += 1
    assert k == int(g['n_params'])
    assert len(net.tr_layers) == len(net.te_layers) == len(c['layers'])
    # the test twins share the parameter objects of the train twins (weights.py:70-76)
    for tr, te in zip(net.tr_layers, net.te_layers):
        assert all(a is b for a, b in zip(tr.params, te.params))


def test_learning_rate_schedule_and_epoch_counter():
    """neuralnet.py:303-314: lr = INIT / (1 + epoch / HALF); the counter lives in training_params
    (and therefore in the .pkl)."""
    c = MR.CASES['plain']
    layers, tp = copy.deepcopy(c['layers']), copy.deepcopy(c['tp'])
    net = nn.NeuralNet(layers, tp, device='cpu')
    assert tp['CUR_EPOCH'] == 0 and net.get_epoch() == 0
    for e in range(4):
        want = np.float32(tp['INIT_LEARNING_RATE'] / (1 + e / tp['EPOCHS_TO_HALF_RATE']))
        assert np.float32(net.cur_learn_rate.get_value()) == want
        net.inc_epoch_set_rate()
    assert tp['CUR_EPOCH'] == 4 and net.get_init_params()['training_params'] is tp
    assert net.get_init_params()['layers'] is layers          # aliasing preserved


def test_weights_round_trip_through_allwts():
    c, g, net = build('mixed')
    saved = net.get_init_params()
    net2 = nn.NeuralNet(saved['layers'], saved['training_params'], saved['allwts'], device='cpu')
    assert net2.rand_gen is None
    for a, b in zip(saved['allwts'], net2.get_init_params()['allwts']):
        assert len(a) == len(b)
        for t, u in zip(a, b):
            assert np.array_equal(t, u)
    assert [len(w) for w in saved['allwts']] == [0, 2, 0, 2, 0, 0, 2, 2, 2]


def test_error_behaviour_follows_the_reference():
    tp = {'SEED': 1, 'BATCH_SZ': 4, 'INIT_LEARNING_RATE': .1, 'EPOCHS_TO_HALF_RATE': 1}
    inp = ('InputLayer', {'img_sz': 8})
    soft = ('SoftmaxLayer', {'n_out': 3})
    with pytest.raises(AssertionError, match="First layer needs to be"):          # neuralnet.py:88-89
        nn.NeuralNet([('ConvLayer', {'num_maps': 2, 'filter_sz': 3, 'stride': 1}), soft], dict(tp), device='cpu')
    with pytest.raises(NotImplementedError, match="Unknown Layer Type"):          # neuralnet.py:196
        nn.NeuralNet([inp, ('NoSuchLayer', {}), soft], dict(tp), device='cpu')
    with pytest.raises(NotImplementedError, match="Unknown Activation Specified"):  # layer.py:54
        nn.NeuralNet([inp, ('HiddenLayer', {'n_out': 4, 'actvn': 'swish'}), soft], dict(tp), device='cpu')
    with pytest.raises(NotImplementedError, match="Loss : "):                     # outlayers.py:36
        nn.NeuralNet([inp, ('SoftmaxLayer', {'n_out': 3, 'loss': 'l2'})], dict(tp), device='cpu')
    with pytest.raises(AssertionError):                                           # convpool.py:40
        nn.NeuralNet([inp, ('ConvLayer', {'num_maps': 2, 'filter_sz': 3, 'stride': 1, 'mode': 'circular'}),
                      soft], dict(tp), device='cpu')
    with pytest.raises(AssertionError):                                           # inlayers.py:66
        nn.NeuralNet([('ElasticLayer', {'img_sz': 8, 'zoom': 0}), soft], dict(tp), device='cpu')
    with pytest.raises(AssertionError):                                           # color.py:30
        nn.NeuralNet([('ColorLayer', {'img_sz': 8, 'gamma': -1}), soft], dict(tp), device='cpu')
    with pytest.raises(KeyError):                                                 # neuralnet.py:66
        nn.NeuralNet([inp, soft], {'BATCH_SZ': 4}, device='cpu')
    # an unreadable truncation threshold falls back to plain NLL (outlayers.py:20-27)
    net = nn.NeuralNet([inp, ('SoftmaxLayer', {'n_out': 3, 'loss': 'nllxy'})], dict(tp), device='cpu')
    assert net.log_thr == 0.0


def test_there_is_no_cpu_execution_path():
    c, g, net = build('plain')
    f = net.get_trin_model(g['x'], g['y'])
    with pytest.raises(RuntimeError, match="no CPU execution path"):
        f(0)
    with pytest.raises(RuntimeError, match="no CPU execution path"):
        net.get_test_model(g['x'], g['y'])(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module under theanet_b200/ (nor train.py) may import
    it, the shim or the reference."""
    import re
    bad = re.compile(r'^\s*(from|import)\s+(oracle|theano|theanet)\b', re.M)
    files = [os.path.join(ROOT, 'train.py')]
    for d, _, fs in os.walk(os.path.join(ROOT, 'theanet_b200')):
        files += [os.path.join(d, f) for f in fs if f.endswith('.py')]
    for f in files:
        assert not bad.search(open(f).read()), f


def test_train_driver_helpers():
    """train.py: rotating test windows (reference train.py:170-176), epoch batch order, image
    reshaping (reference train.py:22-34)."""
    import train
    w = train.window_indices(1000, 100, 300)
    assert [next(w) for _ in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 0, 1]]
    assert train.epoch_batches(105, 10) == list(range(10))
    rs = np.random.RandomState(3)
    b = train.epoch_batches(105, 10, rs)
    flat = np.concatenate(b)
    assert len(b) == 10 and all(v.dtype == np.int32 and len(v) == 10 for v in b)
    assert len(set(flat.tolist())) == 100 and flat.max() < 105
    assert train.as_images(np.zeros((5, 64))).shape == (5, 1, 8, 8)
    assert train.as_images(np.zeros((5, 8, 8))).shape == (5, 1, 8, 8)
    assert train.as_images(np.zeros((5, 3, 8, 8))).shape == (5, 3, 8, 8)
    with pytest.raises(ValueError):
        train.as_images(np.zeros((5, 63)))


def test_resume_state_carries_every_stream_seed():
    """get_resume_state / set_resume_state (an extension over the reference's .pkl): momentum
    buffers, step counter and the seeds of every random stream -- dropout, elastic, colour and the
    LocationInfo mix -- so that a net rebuilt from `allwts` continues the same streams."""
    for name in ('softaux', 'mnist', 'color_el'):
        c = MR.CASES[name]
        net = nn.NeuralNet(copy.deepcopy(c['layers']), copy.deepcopy(c['tp']), device='cpu')
        net.step_count = 7
        st = net.get_resume_state()
        saved = net.get_init_params()
        net2 = nn.NeuralNet(saved['layers'], saved['training_params'], saved['allwts'], device='cpu')
        net2.set_resume_state(st)
        st2 = net2.get_resume_state()
        assert st2['seeds'] == st['seeds'] and st2['aux_seeds'] == st['aux_seeds']
        assert st2['step_count'] == 7
        assert any(s is not None for s in st['seeds'] + st['aux_seeds'])
        for a, b in zip(st['velocities'], st2['velocities']):
            assert all(np.array_equal(u, v) for u, v in zip(a, b))
