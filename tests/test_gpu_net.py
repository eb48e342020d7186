"""Whole-step parity: NeuralNet (CUDA kernels through the C ABI) against the CPU oracle over many
training steps on the same seeded data, with the SAME Philox streams on both sides (the oracle
regenerates the device's dropout masks, flip noise and elastic field bit for bit).

Tolerance (north_star): max|a-b| / max|b| <= 1e-3 per tensor in float32; the first step must leave
every parameter bit-identical (lagged momentum).

The synthetic images here are dense (no exactly-equal neighbouring pixels): max-pool ties route
the gradient to every tied element, so on flat regions a last-ulp difference in a conv sum
between two correct implementations changes which elements are "tied" -- the per-kernel tests
cover tie routing bit-exactly on identical inputs instead."""
import ast
import copy
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import theanet_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def load_prms(name, B, img_sz, seed=555555):
    with open(os.path.join(ROOT, 'params', name)) as f:
        p = ast.literal_eval(f.read())
    p['training_params']['SEED'] = seed
    p['training_params']['BATCH_SZ'] = B
    p['layers'][0][1]['img_sz'] = img_sz
    return p


def synth(n, c, s, n_classes, seed=1234, dense=True):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, c, s, s)).astype(np.float32)
    if not dense:
        x = x * (x > .8)
    y = rng.integers(0, n_classes, n).astype(np.int32)
    return x, y


SMALL_NET = {
    "layers": [
        ('ElasticLayer', {'num_maps': 3, 'translation': 1.5, 'zoom': 1.2, 'magnitude': 12,
                          'sigma': 3, 'pflip': 0.02, 'angle': 10, 'nearest': False,
                          'invert_image': False}),
        ('ConvLayer', {'num_maps': 8, 'filter_sz': 3, 'stride': 1, 'mode': 'same',
                       'actvn': 'relu05', 'reg': {'maxnorm': .8}}),
        ('ConvLayer', {'num_maps': 6, 'filter_sz': 5, 'stride': 1, 'mode': 'valid',
                       'actvn': 'tanh', 'reg': {'L1': 1e-4, 'momentum': .9}}),
        ('PoolLayer', {'pool_sz': 3}),
        ('DropOutLayer', {'pdrop': .2}),
        ('HiddenLayer', {'n_out': 64, 'pdrop': .3, 'actvn': 'relu10',
                         'reg': {'maxnorm': 1.5, 'L1': 1e-4}}),
        ('DropOutLayer', {'pdrop': .1}),
        ('HiddenLayer', {'n_out': 33, 'actvn': 'scaled_tanh', 'reg': {'rate': .5}}),
        ('SoftmaxLayer', {'n_out': 11, 'reg': {'L2': 1e-3, 'maxnorm': 2.}}),
    ],
    "training_params": {'BATCH_SZ': 16, 'NUM_EPOCHS': 1, 'EPOCHS_TO_TEST': 1, 'TEST_SAMP_SZ': 16,
                        'INIT_LEARNING_RATE': .2, 'EPOCHS_TO_HALF_RATE': 2, 'SEED': 4242},
}


def compare_nets(net, on, tag):
    for li, (a, b) in enumerate(zip(net.get_init_params()['allwts'], on.get_wts())):
        for k, (u, v) in enumerate(zip(a, b)):
            assert rel(u, v) < TOL, '{} layer {} tensor {}: {}'.format(tag, li, k, rel(u, v))
    for li, (vs, L) in enumerate(zip(net.get_velocities(), on.spec)):
        for k, u in enumerate(vs):
            r = rel(u, L['vel'][k])
            assert r < TOL or np.max(np.abs(L['vel'][k])) < 1e-12, \
                '{} velocity layer {} tensor {}: {}'.format(tag, li, k, r)


def relu_hidden_layers(net):
    """HiddenLayers (not the output layer) whose activation has a kink at zero."""
    return [li for li, l in enumerate(net.tr_layers[:-1])
            if type(l).__name__ == 'HiddenLayer' and str(l.actvn).startswith('relu')]


def check_kinks(on, n_elems):
    """The ReLU derivative jumps at z = 0: a pre-activation within rounding of zero may land on
    either side depending on the summation order of x.W (3xTF32 tensor-core product vs BLAS), and ONE
    such element moves a weight-gradient column by a few per cent of a typical entry.  The oracle takes
    the device's side for those elements (OracleNet.kink_source); here: they are a handful, and every
    one of them is within 1e-5 of zero relative to the layer's largest pre-activation."""
    for li, n, zmax_flip, zmax in on.kink_flips:
        assert n <= 2 + 1e-5 * n_elems, (li, n)
        assert zmax_flip <= 1e-5 * zmax, (li, zmax_flip, zmax)
    on.kink_flips = []


def run_pair(prms, x, y, steps, use_graph, check_at=(1, 2, 5), frozen_first=True, **trin_kw):
    from theanet_b200.neuralnet import NeuralNet
    p_dev, p_cpu = copy.deepcopy(prms), copy.deepcopy(prms)
    net = NeuralNet(p_dev['layers'], p_dev['training_params'], use_graph=use_graph)
    on = O.OracleNet(p_cpu['layers'], p_cpu['training_params'])
    B = prms['training_params']['BATCH_SZ']
    fn = net.get_trin_model(x, y, **trin_kw)
    init = net.get_init_params()['allwts']
    nb = len(x) // B
    costs = []
    kinks = relu_hidden_layers(net)
    for s in range(steps):
        i = s % nb
        cost, feats, lp = fn(i)
        on.kink_source = {li: net.out[li].cpu().numpy() for li in kinks}
        ocost, olp = on.train_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B], step=s, sample0=0)
        check_kinks(on, B * 1000)
        assert abs(cost - ocost) <= TOL * abs(ocost), 'step {} cost {} vs {}'.format(s, cost, ocost)
        assert rel(lp, olp) < TOL, 'step {} logprob {}'.format(s, rel(lp, olp))
        assert feats is lp or np.array_equal(feats, lp)
        costs.append((float(cost), float(ocost)))
        if s == 0 and frozen_first:      # lagged momentum: nothing moves on the first step
            for a, b in zip(init, net.get_init_params()['allwts']):
                for u, v in zip(a, b):
                    assert np.array_equal(u, v)
        if s + 1 in check_at or s + 1 == steps:
            compare_nets(net, on, 'after step {}'.format(s + 1))
        if (s + 1) % nb == 0:
            net.inc_epoch_set_rate()
            on.inc_epoch_set_rate()
    return net, on, costs


@pytest.mark.parametrize('B,use_graph', [(20, False), (128, True)])
def test_mnist_prms_training_matches_oracle(B, use_graph):
    prms = load_prms('mnist.prms', B, 28)
    x, y = synth(B * 4, 1, 28, 10)
    run_pair(prms, x, y, 10, use_graph)


def test_mnist_prms_at_the_bench_batch_size():
    """BASELINE configs[1] exactly as bench.py runs it -- 1024 images per step, CUDA graph, tcgen05
    3xTF32 dense layers, fused conv+pool and classifier-head kernels, elastic-field prefetch --
    against the oracle for three steps (the sharding property at this size, gradient of the batch =
    mean of the shard gradients, is covered by tests/test_gpu_dist.py and tests/test_dist_gloo.py)."""
    prms = load_prms('mnist.prms', 1024, 28)
    x, y = synth(2048, 1, 28, 10)
    net, on, costs = run_pair(prms, x, y, 3, True, check_at=(1, 3))
    assert net.head and net.conv_fused and net.field_prefetch      # the bench path, not a fallback
    assert all(np.isfinite(c[0]) for c in costs)


def test_mnist_prms_100_steps_curve():
    prms = load_prms('mnist.prms', 20, 28)
    x, y = synth(20 * 10, 1, 28, 10)
    _, _, costs = run_pair(prms, x, y, 100, True, check_at=(1, 2, 10, 100))
    c = np.array(costs)
    assert np.max(np.abs(c[:, 0] - c[:, 1]) / np.abs(c[:, 1])) < TOL


def test_3flat_prms_training_matches_oracle():
    prms = load_prms('3flat.prms', 20, 28)
    x, y = synth(20 * 3, 1, 28, 457)
    run_pair(prms, x, y, 8, True)


def test_small_net_all_layer_kinds_bilinear_maxnorm_l1():
    prms = copy.deepcopy(SMALL_NET)
    prms['layers'][0][1]['img_sz'] = 17
    x, y = synth(16 * 3, 3, 17, 11)
    # maxnorm rescales the initial filters on step 1, so the 'nothing moves' check is off here
    run_pair(prms, x, y, 12, False, frozen_first=False)
    run_pair(prms, x, y, 12, True, frozen_first=False)


def test_graph_and_eager_are_bit_identical_and_host_streaming_matches():
    prms = load_prms('mnist.prms', 32, 28)
    x, y = synth(32 * 3, 1, 28, 10)
    outs = []
    for kw in (dict(use_graph=False), dict(use_graph=True), dict(use_graph=True, resident=False)):
        from theanet_b200.neuralnet import NeuralNet
        p = copy.deepcopy(prms)
        net = NeuralNet(p['layers'], p['training_params'], use_graph=kw.pop('use_graph'))
        fn = net.get_trin_model(x, y, **kw)
        res = [fn(i % 3) for i in range(7)]
        outs.append((res, net.get_init_params()['allwts']))
    for other in outs[1:]:
        for (c0, f0, _), (c1, f1, _) in zip(outs[0][0], other[0]):
            assert c0 == c1 and np.array_equal(f0, f1)
        for a, b in zip(outs[0][1], other[1]):
            for u, v in zip(a, b):
                assert np.array_equal(u, v)


def test_index_list_mode_matches_slices():
    from theanet_b200.neuralnet import NeuralNet
    prms = load_prms('mnist.prms', 16, 28)
    x, y = synth(64, 1, 28, 10)
    perm = np.random.default_rng(3).permutation(64).astype(np.int32)
    p1, p2 = copy.deepcopy(prms), copy.deepcopy(prms)
    n1 = NeuralNet(p1['layers'], p1['training_params'])
    n2 = NeuralNet(p2['layers'], p2['training_params'])
    f1 = n1.get_trin_model(x, y, take_index_list=True)
    f2 = n2.get_trin_model(x[perm], y[perm])
    for i in range(4):
        c1, l1, _ = f1(perm[i * 16:(i + 1) * 16])
        c2, l2, _ = f2(i)
        assert c1 == c2 and np.array_equal(l1, l2)


def test_test_twin_and_checkpoint_roundtrip(tmp_path):
    from theanet_b200.neuralnet import NeuralNet
    prms = load_prms('mnist.prms', 25, 28)
    x, y = synth(100, 1, 28, 10, dense=False)          # MNIST-like sparse images are fine forward
    p_dev, p_cpu = copy.deepcopy(prms), copy.deepcopy(prms)
    net = NeuralNet(p_dev['layers'], p_dev['training_params'])
    on = O.OracleNet(p_cpu['layers'], p_cpu['training_params'])
    te = net.get_test_model(x, y, preds_feats=True)
    for i in range(4):
        err, pmle, feats, preds = te(i)
        oerr, opmle, olp, opreds = on.test_step(x[i * 25:(i + 1) * 25], y[i * 25:(i + 1) * 25])
        assert rel(feats, olp) < TOL
        assert np.array_equal(preds, opreds) and preds.dtype == np.int64
        assert abs(err - oerr) < 1e-6 and abs(pmle - opmle) < TOL * abs(opmle)
    # train a little, checkpoint through pickle (the reference's .pkl schema), reload, same outputs
    tr = net.get_trin_model(x, y)
    for i in range(4):
        tr(i)
    pkl = tmp_path / 'net.pkl'
    with open(pkl, 'wb') as f:
        pickle.dump(net.get_init_params(), f, -1)
    with open(pkl, 'rb') as f:
        saved = pickle.load(f)
    assert [len(w) for w in saved['allwts']] == [0, 2, 0, 2, 0, 2, 2]
    assert saved['allwts'][1][0].shape == (4, 1, 3, 3) and saved['allwts'][5][0].shape == (720, 500)
    net2 = NeuralNet(saved['layers'], saved['training_params'], saved['allwts'])
    te2 = net2.get_test_model(x, y, preds_feats=True)
    a, b = te(1), te2(1)
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])
    dt = net2.get_data_test_model(get_output_of_layers=(2,))
    feats, preds, pooled = dt(x[25:50])
    assert np.array_equal(feats, b[2]) and np.array_equal(preds, b[3])
    assert pooled.shape == (25, 4, 13, 13)


def test_elastic_debugout_matches_oracle_field():
    from theanet_b200.neuralnet import NeuralNet
    from oracle import philox
    prms = load_prms('mnist.prms', 8, 28)
    x, y = synth(16, 1, 28, 10)
    net = NeuralNet(prms['layers'], prms['training_params'], use_graph=False)
    net.debug_elastic = True
    fn = net.get_trin_model(x, y)
    fn(0)
    fn(1)
    seed = net.tr_layers[0].seed
    noise = philox.elastic_noise(seed, 1, 2 * 28 * 28).reshape(2, 28, 28)
    u = philox.elastic_scalars(seed, 1)
    ty, tx, disp = O.elastic_target(28, prms['layers'][0][1], noise, u)
    got_disp, got_tyx = net.elastic_debugout()
    assert np.max(np.abs(got_disp - disp)) < 1e-4        # noise may differ in the last float32 ulp
    assert np.max(np.abs(got_tyx[0] - ty)) < 1e-4
    # the output of the layer is the oracle's warp of the same batch, bit for bit
    fm = philox.bernoulli_mask(seed, philox.PURPOSE_FLIP, 1, np.arange(8), 784, .03).reshape(8, 1, 28, 28)
    want = O.elastic_apply(x[8:16], prms['layers'][0][1], ty, tx, fm)
    assert np.mean(net.out[0].cpu().numpy() == want) > 0.999


# --------------------------------------------------------------------------------------------
# config C4: CIFAR-shaped 3-conv network, bf16 tensor-core conv stack
# --------------------------------------------------------------------------------------------
TOL_BF16 = 1e-2     # bf16 conv stack vs the restatement that rounds the same tensors to bf16: what
                    # is left is accumulation order plus values that straddle a bf16 rounding
                    # boundary (one bf16 ulp = 4e-3 relative on single elements)


def rel2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_cifar3conv_bf16_tensor_core_stack_matches_oracle():
    from theanet_b200.neuralnet import NeuralNet
    B = 8
    prms = load_prms('cifar3conv.prms', B, 32, seed=31337)
    prms['training_params']['CONV_DTYPE'] = 'bfloat16'
    x, y = synth(2 * B, 3, 32, 10, dense=True)
    p1, p2 = copy.deepcopy(prms), copy.deepcopy(prms)
    net = NeuralNet(p1['layers'], p1['training_params'])
    assert sorted(net.conv_tc) == [1, 3, 5]           # all three convs on tcgen05 (conv 1 via im2col)
    assert net.conv_tc[1].im2col and not net.conv_tc[1].fuse_pool and net.conv_tc[3].fuse_pool
    on = O.OracleNet(p2['layers'], p2['training_params'])
    assert [L['tc'] for L in on.spec if L['kind'] == 'ConvLayer'] == [True, True, True]
    fn = net.get_trin_model(x, y)
    for s in range(3):
        i = s % 2
        cost, _, lp = fn(i)
        ocost, olp = on.train_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B], step=s, sample0=0)
        tol = TOL_BF16 if s < 2 else 3 * TOL_BF16
        assert abs(cost - ocost) <= tol * abs(ocost), (s, cost, ocost)
        assert rel(lp, olp) < tol, (s, rel(lp, olp))
        if s >= 2:
            # lagged momentum: steps 0 and 1 run on identical weights; from step 2 on the two
            # sides train on weights that differ in the last bits, and bf16 rounding of the
            # activations amplifies that (a different path through the same noise, not an error)
            continue
        # gradients in the relative L2 norm: an activation that straddles a bf16 rounding boundary
        # flips a max-pool tie and moves single gradient entries by whole contributions (at B = 8
        # that is percents of the largest entry), without touching the bulk of the tensor
        for li, (gg, og) in enumerate(zip(net.get_gradients(), on.last_grads)):
            for k, u in enumerate(gg):
                assert rel2(u, og[k]) < TOL_BF16, 'step {} grad layer {} tensor {}: {}'.format(
                    s, li, k, rel2(u, og[k]))
    for a, b in zip(net.get_init_params()['allwts'], on.get_wts()):
        for u, v in zip(a, b):
            assert rel2(u, v) < TOL_BF16
    e, pr = net.get_test_model(x, y)(0)
    oe, opr, _, _ = on.test_step(x[:B], y[:B])
    assert abs(e - oe) < 1e-6 and abs(pr - opr) <= TOL_BF16 * opr


BF16_NO_POOL = {
    'hidden': [('ConvLayer', {'num_maps': 64, 'filter_sz': 3, 'stride': 1, 'mode': 'same', 'actvn': 'relu50'}),
               ('HiddenLayer', {'n_out': 64, 'pdrop': .25, 'actvn': 'relu10'})],
    'dropout': [('ConvLayer', {'num_maps': 64, 'filter_sz': 3, 'stride': 1, 'mode': 'same', 'actvn': 'relu50'}),
                ('DropOutLayer', {'pdrop': .3}),
                ('HiddenLayer', {'n_out': 64, 'pdrop': 0, 'actvn': 'relu10'})],
}


@pytest.mark.parametrize('case', sorted(BF16_NO_POOL))
def test_bf16_conv_without_pool_feeding_float32_consumers(case):
    """A tensor-core ConvLayer with a leaky reluNN and NO PoolLayer above it, feeding a HiddenLayer /
    a DropOutLayer: its act' must be applied exactly once (by the conv branch of the backward pass,
    not again by the consumer's fused epilogue -- the negative-side slope would come out squared)."""
    from theanet_b200.neuralnet import NeuralNet
    B = 8
    layers = [('InputLayer', {'img_sz': 8, 'num_maps': 64})] + BF16_NO_POOL[case] + \
        [('SoftmaxLayer', {'n_out': 10})]
    tp = {'BATCH_SZ': B, 'NUM_EPOCHS': 1, 'EPOCHS_TO_TEST': 1, 'TEST_SAMP_SZ': B, 'INIT_LEARNING_RATE': .05,
          'EPOCHS_TO_HALF_RATE': 2, 'SEED': 99, 'CONV_DTYPE': 'bfloat16'}
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2 * B, 64, 8, 8)).astype(np.float32)        # half of the pre-activations < 0
    y = rng.integers(0, 10, 2 * B).astype(np.int32)
    net = NeuralNet(copy.deepcopy(layers), dict(tp))
    assert sorted(net.conv_tc) == [1] and net.conv_tc[1].pool is None
    on = O.OracleNet(copy.deepcopy(layers), dict(tp))
    fn = net.get_trin_model(x, y)
    for s in range(2):
        cost, _, lp = fn(s)
        ocost, olp = on.train_step(x[s * B:(s + 1) * B], y[s * B:(s + 1) * B], step=s, sample0=0)
        assert abs(cost - ocost) <= TOL_BF16 * abs(ocost), (s, cost, ocost)
        for li, (gg, og) in enumerate(zip(net.get_gradients(), on.last_grads)):
            for k, u in enumerate(gg):
                assert rel2(u, og[k]) < TOL_BF16, 'step {} grad layer {} tensor {}: {}'.format(
                    s, li, k, rel2(u, og[k]))


@pytest.mark.parametrize('use_graph', [True, False])
def test_exact_resume_with_momentum_step_and_stream_seeds(tmp_path, use_graph):
    """The reference's .pkl carries weights only: a resumed run starts with zero momentum and new
    random streams.  get_resume_state()/set_resume_state() add the momentum
    buffers, the step counter and the stream seeds: 3 + 3 steps across a pickle round trip are
    bit-identical to 6 steps in one go (elastic field, flip noise and dropout masks included)."""
    from theanet_b200.neuralnet import NeuralNet
    prms = load_prms('mnist.prms', 16, 28)
    x, y = synth(64, 1, 28, 10)
    p1, p2 = copy.deepcopy(prms), copy.deepcopy(prms)
    ref = NeuralNet(p1['layers'], p1['training_params'], use_graph=use_graph)
    f_ref = ref.get_trin_model(x, y)
    want = [f_ref(i % 4) for i in range(6)]
    net = NeuralNet(p2['layers'], p2['training_params'], use_graph=use_graph)
    f = net.get_trin_model(x, y)
    for i in range(3):
        f(i % 4)
    pkl = tmp_path / 'resume.pkl'
    with open(pkl, 'wb') as fh:
        pickle.dump(dict(net.get_init_params(), resume=net.get_resume_state()), fh, -1)
    with open(pkl, 'rb') as fh:
        saved = pickle.load(fh)
    net2 = NeuralNet(saved['layers'], saved['training_params'], saved['allwts'], use_graph=use_graph)
    net2.set_resume_state(saved['resume'])
    f2 = net2.get_trin_model(x, y)
    for i in range(3, 6):
        cost, feats, lp = f2(i % 4)
        assert cost == want[i][0] and np.array_equal(lp, want[i][2])
    for a, b in zip(net2.get_init_params()['allwts'], ref.get_init_params()['allwts']):
        for t, u in zip(a, b):
            assert np.array_equal(t, u)
    # without the resume state the weights carry over but momentum and streams do not
    net3 = NeuralNet(saved['layers'], saved['training_params'], saved['allwts'], use_graph=use_graph)
    cost3 = net3.get_trin_model(x, y)(3)[0]
    assert cost3 != want[3][0]


def test_field_prefetch_survives_alternating_corpora_and_test_calls():
    """Under CUDA graphs the elastic field of step s+1 is computed on a side branch of step s
    (NeuralNet._train_step).  Alternating two compiled training functions, test calls in between
    and a learning-rate change must give the eager sequence bit for bit."""
    from theanet_b200.neuralnet import NeuralNet
    prms = load_prms('mnist.prms', 16, 28)
    xa, ya = synth(64, 1, 28, 10, seed=1)
    xb, yb = synth(64, 1, 28, 10, seed=2)

    def run(use_graph):
        p = copy.deepcopy(prms)
        net = NeuralNet(p['layers'], p['training_params'], use_graph=use_graph)
        assert net.field_prefetch == use_graph
        fa, fb = net.get_trin_model(xa, ya), net.get_trin_model(xb, yb)
        te = net.get_test_model(xb, yb)
        out = []
        for s in range(7):
            f = fa if s % 3 else fb
            out.append(f(s % 4)[0])
            if s == 2:
                out.append(te(1)[1])
            if s == 4:
                net.inc_epoch_set_rate()
        return out, net.get_init_params()['allwts']

    (c_g, w_g), (c_e, w_e) = run(True), run(False)
    assert c_g == c_e
    for a, b in zip(w_g, w_e):
        for t, u in zip(a, b):
            assert np.array_equal(t, u)


# --------------------------------------------------------------------------------------------
# lazy returns: the host runs ahead of the device
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize('index_list', [False, True])
def test_lazy_steps_equal_synchronous_steps(index_list):
    """lazy=True returns device tensors without synchronising, so the host enqueues many steps
    ahead.  Per-step scalars travel as launch arguments (tn_set_ctl) and index vectors through an
    event-guarded ring of pinned buffers, so N lazy steps equal N synchronous steps bit for bit
    (a pinned control block read by an in-graph copy could be overwritten by a later step)."""
    from theanet_b200.neuralnet import NeuralNet
    B, N = 64, 40
    prms = load_prms('mnist.prms', B, 28)
    x, y = synth(B * 8, 1, 28, 10)
    order = np.random.default_rng(9).permutation(8)
    perm = np.random.default_rng(10).permutation(B * 8).astype(np.int32)

    def arg(s):
        i = int(order[s % 8])
        return perm[i * B:(i + 1) * B] if index_list else i

    outs = []
    for lazy in (False, True):
        p = copy.deepcopy(prms)
        net = NeuralNet(p['layers'], p['training_params'])
        fn = net.get_trin_model(x, y, take_index_list=index_list, lazy=lazy)
        costs = []
        for s in range(N):
            if s == N // 2:
                net.inc_epoch_set_rate()          # a learning-rate change in mid-flight
            c, _, lp = fn(arg(s))
            costs.append(c.clone() if lazy else c)  # lazy: the device scalar of THAT step
        torch.cuda.synchronize()
        costs = [float(c.item()) if lazy else float(c) for c in costs]
        outs.append((costs, net.get_init_params()['allwts'], net.get_velocities()))
    assert outs[0][0] == outs[1][0]
    for k in (1, 2):
        for a, b in zip(outs[0][k], outs[1][k]):
            for u, v in zip(a, b):
                assert np.array_equal(u, v)


# --------------------------------------------------------------------------------------------
# the benchmark's data regime: 80 % exactly-zero pixels, max-pool ties everywhere
# --------------------------------------------------------------------------------------------
def tie_windows(a_dev, a_ref, p=2):
    """Number of pool windows whose tie pattern (set of elements equal to the window max)
    differs between two activation tensors, and the number of windows."""
    def hits(a):
        out, (xp, o, S, pp, n) = O.pool_forward(a, p, False)
        B, C = xp.shape[:2]
        return xp.reshape(B, C, n, p, n, p) == o[:, :, :, None, :, None]
    ha, hb = hits(a_dev), hits(a_ref)
    diff = (ha != hb).any(axis=(3, 5))
    return int(diff.sum()), int(diff.size)


def run_tie_localised(prms, x, y, steps, pools, **net_kw):
    """Train on data full of exact ties.  Which mathematically equal convolution sums stay equal in
    float32 depends on the summation order, so the CUDA kernels and the NumPy oracle may route the
    max-pool gradient of a few windows differently (assumption A3; no order is canonical).  The
    oracle is therefore told, per step, the tie pattern of the DEVICE's activations
    (OracleNet.tie_source): with that, everything must agree to 1e-3 -- i.e. the whole difference
    between the two implementations is confined to those windows.  Returns the window counts."""
    from theanet_b200.neuralnet import NeuralNet
    p_dev, p_cpu = copy.deepcopy(prms), copy.deepcopy(prms)
    net = NeuralNet(p_dev['layers'], p_dev['training_params'], **net_kw)
    net.keep_conv_out = True        # the comparison below reads the un-pooled conv outputs
    on = O.OracleNet(p_cpu['layers'], p_cpu['training_params'])
    B = prms['training_params']['BATCH_SZ']
    fn = net.get_trin_model(x, y)
    nb = len(x) // B
    ndiff = nwin = 0
    for s in range(steps):
        i = s % nb
        cost, _, lp = fn(i)
        acts = {li: net.out[li - 1].cpu().numpy() for li in pools}
        on.tie_source = acts
        ocost, olp = on.train_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B], step=s, sample0=0)
        assert abs(cost - ocost) <= TOL * abs(ocost), (s, cost, ocost)
        assert rel(lp, olp) < TOL, (s, rel(lp, olp))
        for li in pools:
            a_ref = on.last_caches[li - 1]['a']
            assert rel(acts[li], a_ref) < 1e-5, (s, li)
            d, n = tie_windows(acts[li], a_ref)
            ndiff, nwin = ndiff + d, nwin + n
        for li, (gg, og) in enumerate(zip(net.get_gradients(), on.last_grads)):
            for k, u in enumerate(gg):
                assert rel(u, og[k]) < TOL, 'step {} grad layer {} tensor {}: {}'.format(s, li, k, rel(u, og[k]))
    compare_nets(net, on, 'tie-localised, after {} steps'.format(steps))
    return net, ndiff, nwin


def test_bench_corpus_at_the_bench_batch_size_tie_localised():
    """BASELINE configs[1] on bench.py's own corpus (pixels thresholded so that 80 %
    are exactly 0), 1024 images per step, CUDA graph: the regime in which the tie path of the
    fused conv+pool backward kernels does real work."""
    import bench
    prms = bench.load_prms(1024)
    x, y = bench.synth_corpus(3 * 1024)
    net, ndiff, nwin = run_tie_localised(prms, x, y, 3, pools=(2, 4))
    assert net.head and net.conv_small and net.field_prefetch          # the bench path
    print('tie pattern differs in {} of {} pool windows'.format(ndiff, nwin))
    assert ndiff <= 1e-3 * nwin


def test_c5_shape_64x64_matches_oracle():
    """BASELINE configs[4]: mnist.prms topology on 64x64 images (dense 4500 -> 500), elastic
    distortion on, 512 images per GPU (8 x 512 = the 4096 global batch), three steps."""
    prms = load_prms('mnist.prms', 512, 64)
    x, y = synth(1024, 1, 64, 10)
    net, on, _ = run_pair(prms, x, y, 3, True, check_at=(1, 3))
    assert net.conv_small and net.head


def test_3flat_prms_at_the_bench_batch_size():
    """BASELINE configs[2]: params/3flat.prms (784 -> 1000 -> 457, no conv) at 1024 images per step."""
    prms = load_prms('3flat.prms', 1024, 28)
    x, y = synth(2048, 1, 28, 457)
    run_pair(prms, x, y, 3, True, check_at=(1, 3))


def test_batch_index_bounds_and_graph_lifetime():
    """a batch index past the corpus must raise (the reference's Theano slice would) instead
    of reading out of bounds, and the captured graphs of a closure die with it (they bake the raw
    pointers of ITS staged corpus; a recycled id() must never find them)."""
    import gc
    from theanet_b200.neuralnet import NeuralNet
    prms = load_prms('mnist.prms', 16, 28)
    x, y = synth(48, 1, 28, 10)
    net = NeuralNet(prms['layers'], prms['training_params'])
    fn = net.get_trin_model(x, y)
    te = net.get_test_model(x, y)
    fn(0), fn(2), te(2)
    with pytest.raises(IndexError):
        fn(3)
    with pytest.raises(IndexError):
        te(3)
    with pytest.raises(IndexError):
        fn(-1)
    fn_idx = net.get_trin_model(x, y, take_index_list=True)
    fn_idx(np.arange(16))
    with pytest.raises(IndexError):
        fn_idx(np.arange(40, 56))
    n_graphs = len(net._graphs)
    assert n_graphs >= 3, sorted(net._graphs)
    del fn, te, fn_idx
    gc.collect()
    gc.collect()
    assert len(net._graphs) == 0, sorted(net._graphs)
    fn2 = net.get_trin_model(x[:32], y[:32])
    cost, _, _ = fn2(1)
    assert np.isfinite(cost)
