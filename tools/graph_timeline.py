#!/usr/bin/env python
"""Warm, in-situ timeline of the captured training step: torch.profiler (CUPTI activity records, no
replay, no cache flush) over N graph replays of the bench workload; per kernel of a step its mean
start offset, duration and stream, and the idle gaps on the critical path.

    python tools/graph_timeline.py [--cfg c2] [--steps 30] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys
from collections import defaultdict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='c2')
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    from theanet_b200.neuralnet import NeuralNet, DistContext
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    ctx = DistContext()
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    if world > 1:                                   # under torchrun: the data-parallel step
        dist.init_process_group('nccl', device_id=dev)
        ctx = DistContext(rank, world, None)
    c = bench.CONFIGS[a.cfg]
    prms = bench.load_prms(c['per_gpu'] * world, a.cfg)
    x, y = bench.synth_corpus(8 * c['per_gpu'] * world, a.cfg)
    net = NeuralNet(prms['layers'], prms['training_params'], device=dev, dist=ctx)
    fn = net.get_trin_model(x, y, lazy=True)
    for s in range(20):
        fn(s % 8)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for s in range(a.steps):
            fn(s % 8)
        torch.cuda.synchronize()
    if rank != 0:
        dist.barrier()
        os._exit(0)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
           and 'memcpy' not in e.name.lower() and 'memset' not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at the per-step control kernel
    steps, cur = [], []
    for e in evs:
        if 'set_ctl_kernel' in e.name and cur:
            steps.append(cur)
            cur = []
        cur.append(e)
    steps.append(cur)
    steps = [s for s in steps if len(s) == len(steps[len(steps) // 2])][2:]
    n = len(steps[0])
    rows = []
    for i in range(n):
        t0s = [s[0].time_range.start for s in steps]
        st = np.array([s[i].time_range.start - t0 for s, t0 in zip(steps, t0s)], float)
        du = np.array([s[i].time_range.end - s[i].time_range.start for s in steps], float)
        name = steps[0][i].name.split('(')[0].replace('void ', '')
        rows.append({'i': i, 'kernel': name, 'start_us': float(st.mean()), 'dur_us': float(du.mean()),
                     'end_us': float((st + du).mean())})
    period = float(np.mean(np.diff([s[0].time_range.start for s in steps])))
    print('{} steps of {} kernels; step period {:.1f} us'.format(len(steps), n, period))
    print('{:>3} {:>9} {:>8} {:>9}  {}'.format('#', 'start us', 'dur us', 'end us', 'kernel'))
    for r in rows:
        print('{:>3} {:>9.1f} {:>8.1f} {:>9.1f}  {}'.format(r['i'], r['start_us'], r['dur_us'], r['end_us'],
                                                         r['kernel'][:70]))
    busy = sum(r['dur_us'] for r in rows)
    print('sum of kernel durations {:.1f} us; last kernel ends at {:.1f} us'.format(
        busy, max(r['end_us'] for r in rows)))
    if a.out:
        with open(a.out, 'w') as f:
            json.dump({'cfg': a.cfg, 'world': world, 'period_us': period, 'kernels': rows}, f, indent=1)
    if world > 1:
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
