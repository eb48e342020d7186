"""Golden vectors from the REFERENCE'S OWN CODE -- generator (build container only).

Imports rakeshvar/theanet read-only from /root/reference and runs its unmodified NeuralNet /
ElasticLayer over oracle/theano_shim (a torch-backed stand-in for the Theano calls that path makes;
Theano itself cannot be installed here).  For every case it records the inputs, every random draw
the reference's RandomStreams made, and the outputs (cost, log-probabilities, weights and momentum
buffers after the updates, test-model statistics, the elastic layer's image and displacement
field) into tests/golden/ref_<case>.npz.  tests/test_golden_ref.py then feeds the same inputs and
draws to oracle/theanet_oracle.py and compares.

    python tests/golden/make_golden_ref.py            # rewrites tests/golden/ref_*.npz

What this pins and what it cannot: see oracle/theano_shim/README.md.
Nothing on the GPU box runs this file (/root/reference does not exist there); importing it for
CASES / draw_keys is safe anywhere.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = '/root/reference'

REG = {'L1': 1e-4, 'L2': 1e-3, 'momentum': .8, 'maxnorm': 1.5, 'rate': .7}

# name -> (layers, training_params, image size, channels, classes, corpus batches, steps)
CASES = {
    # params/mnist.prms as shipped by the reference (mnist.prms:3-51): nearest-neighbour elastic
    # layer with every distortion switched on, two conv+pool stages, dropout-500 hidden, softmax.
    'mnist': dict(
        layers=[
            ('ElasticLayer', {'img_sz': 28, 'translation': 2, 'zoom': 1.1, 'magnitude': 60, 'sigma': 15,
                              'pflip': 0.03, 'angle': 5, 'nearest': True, 'invert_image': True}),
            ('ConvLayer', {'num_maps': 4, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu10'}),
            ('PoolLayer', {'pool_sz': 2}),
            ('ConvLayer', {'num_maps': 20, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu05'}),
            ('PoolLayer', {'pool_sz': 2}),
            ('HiddenLayer', {'n_out': 500, 'pdrop': .5, 'reg': {'L2': .0, 'maxnorm': 0}}),
            ('SoftmaxLayer', {'n_out': 10, 'reg': {'L2': .0, 'maxnorm': 0}}),
        ],
        tp={'SEED': 555555, 'BATCH_SZ': 8, 'INIT_LEARNING_RATE': .1, 'EPOCHS_TO_HALF_RATE': 1},
        channels=1, classes=10, batches=2, steps=5, bump_epoch_at=3),
    # everything mnist.prms leaves out: bilinear interpolation, 'same' convolution, ignore_border
    # and ragged (ceil) pooling with window 3, tanh / scaled_tanh / relu, a DropOutLayer, L1 + L2,
    # max-norm on 4-D, 2-D and 1-D parameters, non-default momentum and per-layer rate.
    'mixed': dict(
        layers=[
            ('ElasticLayer', {'img_sz': 15, 'num_maps': 2, 'translation': 1.5, 'zoom': 1.25, 'magnitude': 25,
                              'sigma': 3, 'angle': 12}),
            ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'mode': 'same', 'actvn': 'relu',
                           'reg': REG}),
            ('PoolLayer', {'pool_sz': 2, 'ignore_border': True}),
            ('ConvLayer', {'num_maps': 5, 'filter_sz': 3, 'stride': 1, 'actvn': 'tanh', 'reg': REG}),
            ('PoolLayer', {'pool_sz': 3}),
            ('DropOutLayer', {'pdrop': .25}),
            ('HiddenLayer', {'n_out': 24, 'actvn': 'scaled_tanh', 'reg': REG}),
            ('HiddenLayer', {'n_out': 16, 'pdrop': .3, 'actvn': 'sigmoid', 'reg': REG}),
            ('SoftmaxLayer', {'n_out': 7, 'reg': REG}),
        ],
        tp={'SEED': 4242, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .5, 'EPOCHS_TO_HALF_RATE': 2},
        channels=2, classes=7, batches=2, steps=6, bump_epoch_at=2),
    # no distortion at all (InputLayer), three input channels, softplus + relu50 (the default)
    'plain': dict(
        layers=[
            ('InputLayer', {'img_sz': 10, 'num_maps': 3}),
            ('ConvLayer', {'num_maps': 6, 'filter_sz': 5, 'stride': 1}),
            ('PoolLayer', {'pool_sz': 2}),
            ('HiddenLayer', {'n_out': 20, 'actvn': 'softplus'}),
            ('SoftmaxLayer', {'n_out': 4}),
        ],
        tp={'SEED': 7, 'BATCH_SZ': 5, 'INIT_LEARNING_RATE': .2, 'EPOCHS_TO_HALF_RATE': 5},
        channels=3, classes=4, batches=3, steps=6, bump_epoch_at=4),
}


def _small(out_layer, seed):
    """One conv stage + a tanh hidden layer under the output layer / loss being pinned."""
    return dict(
        layers=[
            ('InputLayer', {'img_sz': 8, 'num_maps': 1}),
            ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu20'}),
            ('PoolLayer', {'pool_sz': 2}),
            ('HiddenLayer', {'n_out': 12, 'actvn': 'tanh', 'reg': {'momentum': .5}}),
            out_layer,
        ],
        tp={'SEED': seed, 'BATCH_SZ': 7, 'INIT_LEARNING_RATE': .3, 'EPOCHS_TO_HALF_RATE': 3},
        channels=1, classes=5, batches=2, steps=5, bump_epoch_at=3)


# the other losses and output layers (outlayers.py:38-64,105-147), MeanLayer, ColorLayer
CASES.update({
    'meanpool': dict(
        layers=[
            ('InputLayer', {'img_sz': 8, 'num_maps': 2}),
            ('ConvLayer', {'num_maps': 4, 'filter_sz': 3, 'stride': 1, 'mode': 'same', 'actvn': 'relu10',
                           'reg': {'momentum': .6}}),
            ('MeanLayer', {}),
            ('HiddenLayer', {'n_out': 6, 'actvn': 'tanh', 'reg': {'momentum': .6}}),
            ('SoftmaxLayer', {'n_out': 3, 'reg': {'momentum': .6}}),
        ],
        tp={'SEED': 21, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .4, 'EPOCHS_TO_HALF_RATE': 2},
        channels=2, classes=3, batches=2, steps=5, bump_epoch_at=2),
    # strided convolutions (convpool.py:54-56,69-70): 14 -> (f3, stride 2) 12/2 = 6 -> (f4, stride 3) 3/3 = 1
    'stride': dict(
        layers=[
            ('InputLayer', {'img_sz': 14, 'num_maps': 2}),
            ('ConvLayer', {'num_maps': 4, 'filter_sz': 3, 'stride': 2, 'actvn': 'relu10',
                           'reg': {'momentum': .6}}),
            ('ConvLayer', {'num_maps': 5, 'filter_sz': 4, 'stride': 3, 'actvn': 'tanh',
                           'reg': {'momentum': .6, 'maxnorm': 1.}}),
            ('HiddenLayer', {'n_out': 10, 'actvn': 'relu', 'reg': {'momentum': .6}}),
            ('SoftmaxLayer', {'n_out': 3, 'reg': {'momentum': .6}}),
        ],
        tp={'SEED': 41, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .3, 'EPOCHS_TO_HALF_RATE': 2},
        channels=2, classes=3, batches=2, steps=5, bump_epoch_at=2),
    # auxiliary inputs (auxiliary.py): features of a (2,2) side input appended to the hidden
    # layer's output (never trained: AuxConcatLayer has no reg) ...
    'auxcat': dict(
        layers=[
            ('InputLayer', {'img_sz': 8, 'num_maps': 1}),
            ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu20',
                           'reg': {'momentum': .6}}),
            ('PoolLayer', {'pool_sz': 2}),
            ('HiddenLayer', {'n_out': 12, 'pdrop': .25, 'actvn': 'tanh', 'reg': {'momentum': .6}}),
            ('AuxConcatLayer', {'n_aux': (5, 9), 'aux_type': 'LocationInfo', 'boost': 2}),
            ('SoftmaxLayer', {'n_out': 4, 'reg': {'momentum': .6}}),
        ],
        tp={'SEED': 31, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .3, 'EPOCHS_TO_HALF_RATE': 2},
        channels=1, classes=4, batches=2, steps=5, bump_epoch_at=2),
    # ... or entering the scores of the softmax layer through a trained cross term
    'softaux': dict(
        layers=[
            ('InputLayer', {'img_sz': 8, 'num_maps': 1}),
            ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu20',
                           'reg': {'momentum': .6}}),
            ('PoolLayer', {'pool_sz': 2}),
            ('HiddenLayer', {'n_out': 12, 'actvn': 'tanh', 'reg': {'momentum': .6}}),
            ('SoftAuxLayer', {'n_out': 4, 'n_aux': (5, 9), 'aux_type': 'LocationInfo',
                              'reg': {'momentum': .6, 'L2': 1e-3, 'maxnorm': 2.}}),
        ],
        tp={'SEED': 32, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .3, 'EPOCHS_TO_HALF_RATE': 2},
        channels=1, classes=4, batches=2, steps=5, bump_epoch_at=2),
    # ColorLayer as the input layer ...
    'color0': dict(
        layers=[
            ('ColorLayer', {'img_sz': 8, 'num_maps': 3, 'balance': 1.5, 'gamma': 1.4}),
            ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu',
                           'reg': {'momentum': .6}}),
            ('PoolLayer', {'pool_sz': 2}),
            ('SoftmaxLayer', {'n_out': 4, 'reg': {'momentum': .6}}),
        ],
        tp={'SEED': 22, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .2, 'EPOCHS_TO_HALF_RATE': 2},
        channels=3, classes=4, batches=2, steps=4, bump_epoch_at=2),
    # ... in front of an ElasticLayer (Elastic past position 0, neuralnet.py:132-142) ...
    'color_el': dict(
        layers=[
            ('ColorLayer', {'img_sz': 9, 'num_maps': 2, 'balance': 1.3, 'gamma': 1.2}),
            ('ElasticLayer', {'translation': 1, 'zoom': 1.2, 'magnitude': 10, 'sigma': 2, 'pflip': .05,
                              'angle': 8}),
            ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu',
                           'reg': {'momentum': .6}}),
            ('PoolLayer', {'pool_sz': 2}),
            ('SoftmaxLayer', {'n_out': 4, 'reg': {'momentum': .6}}),
        ],
        tp={'SEED': 24, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .2, 'EPOCHS_TO_HALF_RATE': 2},
        channels=2, classes=4, batches=2, steps=4, bump_epoch_at=2),
    # ... and behind an ElasticLayer, with a value range of [0, 2] (maxval)
    'color1': dict(
        layers=[
            ('ElasticLayer', {'img_sz': 8, 'num_maps': 3, 'translation': 1}),
            ('ColorLayer', {'balance': 1.2, 'gamma': 1.6, 'maxval': 2}),
            ('ConvLayer', {'num_maps': 3, 'filter_sz': 3, 'stride': 1, 'actvn': 'relu',
                           'reg': {'momentum': .6}}),
            ('PoolLayer', {'pool_sz': 2}),
            ('SoftmaxLayer', {'n_out': 4, 'reg': {'momentum': .6}}),
        ],
        tp={'SEED': 23, 'BATCH_SZ': 6, 'INIT_LEARNING_RATE': .2, 'EPOCHS_TO_HALF_RATE': 2},
        channels=3, classes=4, batches=2, steps=4, bump_epoch_at=2),
    'nllsq': _small(('SoftmaxLayer', {'n_out': 5, 'loss': 'nllsq', 'reg': {'momentum': .5}}), 11),
    # truncated NLL: threshold .2 = chance level for 5 classes, so rows fall on both sides of it
    'nll20': _small(('SoftmaxLayer', {'n_out': 5, 'loss': 'nll20', 'reg': {'momentum': .5}}), 12),
    'exploss': _small(('ExpLossLayer', {'n_out': 5, 'reg': {'momentum': .5, 'L2': 1e-3}}), 13),
    'hinge': _small(('HingeLayer', {'n_out': 5, 'reg': {'momentum': .5, 'maxnorm': 2.}}), 14),
})


def case_data(name):
    """Deterministic corpus: images in [0,1] whose top third is a flat background that reaches the
    first conv layer as exact zeros, so that exact ties inside pooling windows (A3) and
    pre-activations exactly equal to the bias -- 0 for reluNN (A5) -- occur and do not depend on
    the summation order of the convolution (a sum of zeros is exact in any order; a flat non-zero
    background under the reference's binary +-c conv weights is not, and every float32
    implementation then breaks those ties its own way).  The foreground is not quantised."""
    c = CASES[name]
    img = c['layers'][0][1]['img_sz']
    n = c['tp']['BATCH_SZ'] * c['batches']
    rs = np.random.RandomState(sum(map(ord, name)))
    x = rs.rand(n, c['channels'], img, img).astype(np.float32)
    # 'mnist' inverts its input (invert_image): paint the background 1 so that it is 0 after it
    x[:, :, :img // 3, :] = 1 if c['layers'][0][1].get('invert_image', False) else 0
    y = rs.randint(0, c['classes'], n).astype(np.int32)
    return x.astype(np.float32), y


def case_aux(name):
    """Auxiliary corpus (n, 2, 2) for the nets that take one (auxiliary.py:22), else None."""
    c = CASES[name]
    if not any(k in ('AuxConcatLayer', 'SoftAuxLayer') for k, _ in c['layers']):
        return None
    n = c['tp']['BATCH_SZ'] * c['batches']
    return np.random.RandomState(99 + sum(map(ord, name))).rand(n, 2, 2).astype(np.float32)


def draw_keys(layers):
    """[(stream index, serial, layer index, oracle key)] in the order the reference constructs its
    RandomStreams and draws from them (inlayers.py:72-141, dropout.py:9-13, hidden.py:32-33)."""
    out, stream = [], 0
    for li, (kind, a) in enumerate(layers):
        if kind == 'ElasticLayer':
            if not (a.get('magnitude', 0) or a.get('translation', 0) or a.get('pflip', 0)
                    or a.get('angle', 0)) and a.get('zoom', 1) == 1:
                continue
            serial = 0
            wanted = [('translation', a.get('translation', 0)), ('noise', a.get('magnitude', 0)),
                      ('origin', a.get('zoom', 1) - 1 or a.get('angle', 0)),
                      ('zoom', a.get('zoom', 1) - 1), ('angle', a.get('angle', 0)),
                      ('flip', a.get('pflip', 0))]
            for key, on in wanted:
                if on:
                    out.append((stream, serial, li, key))
                    serial += 1
            stream += 1
        elif kind in ('DropOutLayer', 'HiddenLayer') and a.get('pdrop', 0):
            out.append((stream, 0, li, 'mask'))
            stream += 1
        elif kind in ('AuxConcatLayer', 'SoftAuxLayer'):     # LocationInfo's mix (auxiliary.py:25-27)
            out.append((stream, 0, li, 'auxu'))
            stream += 1
        elif kind == 'ColorLayer' and not (a.get('gamma', 1) == 1 and a.get('balance', 1) == 1):
            for k in range(3):                      # pos_rand: balance, gamma, gamma (color.py:36-42)
                out.append((stream, k, li, 'color%d' % k))
            stream += 1
    return out


def rand_table(g, layers, prefix):
    """Rebuild the oracle's injected-randomness dict for one call from the arrays stored in g."""
    rand, u, col = {}, {}, {}
    for _, _, li, key in draw_keys(layers):
        v = g['%s_%d_%s' % (prefix, li, key)]
        if key.startswith('color'):
            col.setdefault(li, [None] * 3)[int(key[-1])] = v
        elif key == 'auxu':
            rand[(li, 'auxu')] = v
        elif key in ('flip', 'mask'):
            shape = tuple(g['%s_%d_%s_shape' % (prefix, li, key)])
            rand[(li, key)] = np.unpackbits(v)[:int(np.prod(shape))].reshape(shape).astype(np.float32)
        elif key == 'noise':
            rand[(li, 'noise')] = v
        else:
            u.setdefault(li, {})[key] = v
    for li, d in u.items():
        rand[(li, 'u')] = d
        rand.setdefault((li, 'noise'), None)
    for li, d in col.items():
        rand[(li, 'color')] = np.stack(d, axis=-1)           # (B, maps, 3)
    return rand


def thin(a, limit=16384):
    """Large tensors are stored as a strided sample (the digest in the same file covers the rest)."""
    flat = np.asarray(a).reshape(-1)
    stride = max(1, -(-flat.size // limit))
    return flat[::stride].copy()


def digest(a):
    a = np.asarray(a, np.float64)
    return np.array([a.sum(), (a * a).sum(), np.abs(a).max()])


def _import_reference():
    if not os.path.isdir(REFERENCE):
        raise SystemExit("make_golden_ref.py needs the reference checkout at " + REFERENCE)

    class _Cast:                       # numpy.cast left NumPy in 2.0; neuralnet.py:110 still uses it
        def __getitem__(self, dt):
            return lambda v: np.asarray(v, dtype=dt)
    if not hasattr(np, 'cast'):
        np.cast = _Cast()
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'theano_shim'))
    sys.path.insert(0, REFERENCE)
    import theano
    import theano.tensor as tt
    from theanet.neuralnet import NeuralNet
    return theano, tt, NeuralNet


def store_draws(rec, prefix, draws, layers, base):
    table = {(s, k): (li, key) for s, k, li, key in draw_keys(layers)}
    seen = set()
    for stream, serial, kind, val in draws:
        li, key = table[(stream - base, serial)]
        seen.add((li, key))
        name = '%s_%d_%s' % (prefix, li, key)
        if key in ('flip', 'mask'):
            rec[name] = np.packbits(val.astype(np.uint8).reshape(-1))
            rec[name + '_shape'] = np.array(val.shape)
        else:
            rec[name] = np.asarray(val, np.float32)
    assert len(seen) == len(table), "a random variable of the graph was never evaluated"


def generate(name, write=True):
    theano, tt, NeuralNet = _import_reference()
    from theano.tensor.shared_randomstreams import RandomStreams
    c = CASES[name]
    layers, tp = copy.deepcopy(c['layers']), copy.deepcopy(c['tp'])
    x, y = case_data(name)
    B = tp['BATCH_SZ']
    base = len(RandomStreams.instances)
    net = NeuralNet(layers, tp)                                    # the reference's constructor
    rec = {'x': x, 'y': y}
    # what the reference prints (train.py:115-116,139-140): layer representations of both twins,
    # the layer / training-parameter listings (after the constructor mutated them) and the
    # weight summary -- the drop-in must print the same text
    rec['repr'] = np.array(str(net).split('\nParams')[0])
    rec['layers_info'] = np.array(net.get_layers_info())
    rec['tp_info'] = np.array(net.get_training_params_info())
    rec['wts_info'] = np.array(net.get_wts_info(detailed=True))
    rec['shapes'] = np.array([[getattr(l, 'n_out', -1), getattr(l, 'num_maps', -1) or -1,
                               getattr(l, 'out_sz', -1) or -1] for l in net.tr_layers])
    k = 0
    for lyr in net.tr_layers:
        for p in lyr.params:
            rec['w0_%d' % k] = thin(p.get_value())
            rec['w0d_%d' % k] = digest(p.get_value())
            k += 1
    xs = theano.shared(x, borrow=True)
    ys = tt.cast(theano.shared(y, borrow=True), 'int32')           # train.py:27-31 share()
    aux = case_aux(name)
    auxs = None
    if aux is not None:
        rec['aux'] = aux
        auxs = theano.shared(aux, borrow=True)
    train = net.get_trin_model(xs, ys, auxs)
    test = net.get_test_model(xs, ys, auxs)
    for s in range(c['steps']):
        if s == c['bump_epoch_at']:
            net.inc_epoch_set_rate()
        cost, feats, logprob = train(s % c['batches'])
        rec['cost_%d' % s] = np.float64(cost)
        rec['logprob_%d' % s] = logprob
        rec['feat_%d' % s] = feats
        store_draws(rec, 's%d' % s, train.draws, c['layers'], base)
    k = 0
    for lyr in net.tr_layers:
        accs = getattr(lyr, 'accumulated_updates', [])           # none for layers without reg
        for j, p in enumerate(lyr.params):
            rec['w_%d' % k], rec['wd_%d' % k] = thin(p.get_value()), digest(p.get_value())
            if j < len(accs):
                rec['v_%d' % k], rec['vd_%d' % k] = thin(accs[j].get_value()), digest(accs[j].get_value())
            k += 1
    rec['n_params'] = np.int64(k)
    for b in range(c['batches']):
        err, py = test(b)
        rec['test_%d' % b] = np.array([err, py], np.float64)
    if c['layers'][0][0] == 'ElasticLayer' and 'magnitude' in c['layers'][0][1]:   # tests/test_elastic.py's view
        el = theano.function([net.x], net.tr_layers[0].debugout[:2])
        img, disp = el(x[:B])
        rec['elastic_img'], rec['elastic_disp'] = img, disp
        store_draws(rec, 'el', [d for d in el.draws], c['layers'][:1], base)
    path = os.path.join(HERE, 'ref_%s.npz' % name)
    if not write:
        return rec
    np.savez_compressed(path, **rec)
    print('%-6s -> %s (%.0f KB)  cost %s' % (name, os.path.relpath(path, ROOT), os.path.getsize(path) / 1024,
                                             ' '.join('%.5f' % rec['cost_%d' % s] for s in range(c['steps']))))


def check(name):
    """Re-run the reference for one case and compare with the committed fixture."""
    rec = generate(name, write=False)
    g = np.load(os.path.join(HERE, 'ref_%s.npz' % name))
    assert sorted(rec) == sorted(g.files), "fixture keys differ"
    for k in g.files:
        a, b = np.asarray(rec[k]), g[k]
        assert a.shape == b.shape, k
        if a.dtype.kind == 'f':
            assert np.allclose(a, b, rtol=1e-5, atol=1e-7), k
        else:
            assert np.array_equal(a, b), k
    print('%s: fixture reproduced from %s' % (name, REFERENCE))


if __name__ == '__main__':
    args = sys.argv[1:]
    if args and args[0] == '--check':
        for n in (args[1:] or CASES):
            check(n)
    else:
        for n in (args or CASES):
            generate(n)
