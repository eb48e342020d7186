"""Run a few eager (graph-off) training steps of the bench workload so that ncu sees plain kernel
launches:  ncu ... python tools/profile_step.py [--steps 3] [--batch 1024] [--prms mnist.prms]"""
import argparse
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--prms', default='mnist.prms')
    ap.add_argument('--img', type=int, default=28)
    ap.add_argument('--maps', type=int, default=1)
    ap.add_argument('--classes', type=int, default=10)
    ap.add_argument('--bf16', type=int, default=0)
    ap.add_argument('--graph', type=int, default=0)
    a = ap.parse_args()
    from theanet_b200.neuralnet import NeuralNet
    with open(os.path.join(ROOT, 'params', a.prms)) as f:
        p = ast.literal_eval(f.read())
    p['training_params'].update(SEED=555555, BATCH_SZ=a.batch)
    if a.bf16:
        p['training_params']['CONV_DTYPE'] = 'bfloat16'
    p['layers'][0][1]['img_sz'] = a.img
    rng = np.random.default_rng(1234)
    x = rng.random((a.batch * 4, a.maps, a.img, a.img), dtype=np.float32)
    x *= (x > .8)
    y = rng.integers(0, a.classes, a.batch * 4).astype(np.int32)
    net = NeuralNet(p['layers'], p['training_params'], use_graph=bool(a.graph))
    fn = net.get_trin_model(x, y)
    import time
    import torch
    for s in range(a.steps):
        cost, _, _ = fn(s % 4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(a.steps):
        cost, _, _ = fn(s % 4)
    torch.cuda.synchronize()
    print('ms/step (host-timed, incl. sync + D2H)', 1e3 * (time.perf_counter() - t0) / a.steps)
    print('cost', float(cost), 'launches/step', net.launches.get('train'))


if __name__ == '__main__':
    main()
