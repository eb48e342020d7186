#!/usr/bin/env python
"""Phase timeline of tn_convpool_bwd (clock64 stamps by thread 0 of every CTA) for a layer geometry.

    python tools/phase_times.py [B C S M need_dx]        default: conv 2 of mnist.prms at B=1024
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theanet_b200 import _C as C   # noqa: E402

NAMES = ['start', 'prologue done', 'staged', 'wgrad done', 'db done', 'dgrad done', 'stage barrier',
         'partials written', 'team ticket', 'final']


def main():
    a = [int(v) for v in sys.argv[1:6]] + [1024, 4, 13, 20, 1][len(sys.argv) - 1:]
    B, Cin, S, M, need_dx = a
    O, act, nn = S - 2, *C.act_code('relu05')
    P = (O + 1) // 2
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn((B, Cin, S, S), device='cuda', generator=g)
    W = torch.randn((M, Cin, 3, 3), device='cuda', generator=g) / 6
    b = torch.randn(M, device='cuda', generator=g)
    pooled = torch.zeros((B, M, P, P), device='cuda')
    tie = torch.zeros(B * M * P * P, dtype=torch.uint8, device='cuda')
    dtop = torch.randn((B, M, P, P), device='cuda', generator=g)
    dW, db, dx = torch.zeros_like(W), torch.zeros_like(b), torch.zeros_like(x)
    geom = (Cin, S, M, 3, 0, O, act, 2, P)
    ws = torch.zeros(C.lib.tn_convpool_bwd_workspace_bytes(B, *geom, need_dx) // 4 + 1, device='cuda')
    C.call('tn_convpool_fprop_train', C.ptr(x), C.ptr(W), C.ptr(b), None, C.ptr(pooled), C.ptr(tie), B,
           Cin, S, M, 3, 0, O, act, nn, 2, P, None)
    def fwd():
        C.call('tn_convpool_fprop_train', C.ptr(x), C.ptr(W), C.ptr(b), None, C.ptr(pooled), C.ptr(tie), B,
               Cin, S, M, 3, 0, O, act, nn, 2, P, None)
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(20):
        fwd()
    f1.record()
    torch.cuda.synchronize()
    print('forward kernel: %.2f us per launch (warm L2)' % (1e3 * f0.elapsed_time(f1) / 20))
    dbg = torch.zeros(1024 * 64, dtype=torch.int64, device='cuda')

    def run():
        C.call('tn_convpool_bwd', C.ptr(x), None, C.ptr(tie), C.ptr(pooled), C.ptr(dtop), C.ptr(W),
               C.ptr(dW), C.ptr(db), C.ptr(dx) if need_dx else None, None, C.ptr(ws), B, Cin, S, M, 3,
               0, O, act, nn, 2, P, 0, 0, None)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print('kernel: %.2f us per launch (warm L2)' % (1e3 * e0.elapsed_time(e1) / 20))
    C.call('tn_convpool_debug_timestamps', C.ptr(dbg))
    run()
    torch.cuda.synchronize()
    C.call('tn_convpool_debug_timestamps', None)
    t = dbg.cpu().numpy().reshape(1024, 64)
    t = t[t[:, 0] > 0]
    print('CTAs', len(t))

    def show(name, a, b):
        ok = (a > 0) & (b > 0)
        if ok.any():
            d = (a - b)[ok]
            print('  %-26s n=%4d  cycles: mean %8.0f  max %8.0f' % (name, ok.sum(), d.mean(), d.max()))
    show('prologue (thread 0)', t[:, 1], t[:, 0])
    prev_end = t[:, 1]
    for st in range(6):
        o = st * 8
        if not (t[:, o + 2] > 0).any():
            break
        print(' stage %d%s' % (st, '+' if st == 5 else ''))
        show('staged (incl. barrier)', t[:, o + 2], np.where(t[:, o + 2] > 0, prev_end, 0))
        show('wgrad', t[:, o + 3], t[:, o + 2])
        show('compute (direct wgrad)', t[:, o + 5], np.where(t[:, o + 3] > 0, 0, t[:, o + 2]))
        show('db', t[:, o + 4], t[:, o + 3])
        show('dgrad', t[:, o + 5], t[:, o + 4])
        show('barrier', t[:, o + 6], t[:, o + 5])
        prev_end = t[:, o + 6]
    last = np.max(t[:, [6, 14, 22, 30, 38, 46]], axis=1)
    show('acc -> smem + barrier', t[:, 48], last)
    show('slice tree', t[:, 49], t[:, 48])
    show('partial write', t[:, 50], t[:, 49])
    show('db tree + tail', t[:, 56], t[:, 50])
    show('slice sums + partials', t[:, 56], last)
    show('team ticket', t[:, 57], t[:, 56])
    show('team sum + final (1 CTA)', t[:, 58], t[:, 57])
    show('whole CTA', np.maximum(t[:, 57], t[:, 58]), t[:, 0])


if __name__ == '__main__':
    main()
