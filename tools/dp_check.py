"""Data-parallel parity check on >= 2 GPUs (launch under torchrun): every rank trains its shard,
rank 0 compares cost / log-probabilities / weights with the single-process CPU oracle on the full
minibatch.   torchrun --nproc-per-node 2 tools/dp_check.py [--graph 0|1]"""
import argparse
import ast
import copy
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def log(*a):
    print('[rank %s %.1fs]' % (os.environ.get('RANK', '0'), time.time() - T0), *a, flush=True)


T0 = time.time()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graph', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    a = ap.parse_args()
    from theanet_b200.dist import init_from_env
    from theanet_b200.neuralnet import NeuralNet
    from oracle import theanet_oracle as O
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    ctx = init_from_env(dev)
    log('init done, world', ctx.world)
    with open(os.path.join(ROOT, 'params', 'mnist.prms')) as f:
        p = ast.literal_eval(f.read())
    B = 32 * ctx.world
    p['training_params'].update(SEED=555555, BATCH_SZ=B)
    p['layers'][0][1]['img_sz'] = 28
    rng = np.random.default_rng(3)
    x = rng.random((2 * B, 1, 28, 28), dtype=np.float32)
    x *= (x > .8)
    y = rng.integers(0, 10, 2 * B).astype(np.int32)
    p1, p2 = copy.deepcopy(p), copy.deepcopy(p)
    net = NeuralNet(p1['layers'], p1['training_params'], device=dev, dist=ctx, use_graph=bool(a.graph))
    log('net built')
    fn = net.get_trin_model(x, y)
    on = O.OracleNet(p2['layers'], p2['training_params']) if ctx.rank == 0 else None
    Bl = B // ctx.world
    worst = 0.0
    for s in range(a.steps):
        i = s % 2
        cost, _, lp = fn(i)
        log('step', s, 'cost', float(cost))
        if on is not None:
            ocost, olp = on.train_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B], step=s, sample0=0)
            e1 = abs(float(cost) - float(ocost)) / abs(float(ocost))
            e2 = float(np.max(np.abs(lp - olp[:Bl])) / np.max(np.abs(olp)))
            worst = max(worst, e1, e2)
            log('   vs oracle: cost rel %.2e  logprob(shard) rel %.2e' % (e1, e2))
    if on is not None:
        for u, v in zip([t for l in net.get_init_params()['allwts'] for t in l],
                        [t for l in on.get_wts() for t in l]):
            worst = max(worst, float(np.max(np.abs(u - v)) / max(np.max(np.abs(v)), 1e-30)))
        log('worst relative deviation', '%.2e' % worst)
        assert worst < 1e-3
        print('DP_CHECK_OK world=%d graph=%d worst=%.2e' % (ctx.world, a.graph, worst), flush=True)
    if ctx.world > 1:
        # leave without tearing the communicator down: the CUDA graphs that captured collectives
        # are still alive, and destroy_process_group() behind them has been seen to block
        torch.distributed.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
