"""Determinism stress of tn_convpool_bwd: many repeated launches on identical inputs must give
identical bits.   python tools/stress_convpool_bwd.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theanet_b200 import _C as C  # noqa: E402

CASES = [(700, 4, 13, 20), (5, 2, 33, 5), (1024, 4, 13, 20), (1024, 1, 28, 4), (9, 4, 31, 20)]
if os.environ.get('TN_STRESS_CASE'):
    CASES = [CASES[int(os.environ['TN_STRESS_CASE'])]]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    for (B, Cin, S, M) in CASES:
        O_ = S - 2
        P = (O_ + 1) // 2
        act, nn = C.act_code('relu05')
        g = torch.Generator(device='cuda').manual_seed(1)
        x = torch.randn((B, Cin, S, S), device='cuda', generator=g)
        W = torch.randn((M, Cin, 3, 3), device='cuda', generator=g) / 6
        b = torch.randn(M, device='cuda', generator=g)
        a_full = torch.zeros((B, M, O_, O_), device='cuda')
        pooled = torch.zeros((B, M, P, P), device='cuda')
        tie = torch.zeros(B * M * P * P, dtype=torch.uint8, device='cuda')
        dtop = torch.randn((B, M, P, P), device='cuda', generator=g)
        C.call('tn_convpool_fprop_train', C.ptr(x), C.ptr(W), C.ptr(b), C.ptr(a_full), C.ptr(pooled), C.ptr(tie), B,
               Cin, S, M, 3, 0, O_, act, nn, 2, P, None)
        geom = (Cin, S, M, 3, 0, O_, act, 2, P)
        for need_dx, use_tie, below in ((1, True, False), (0, True, False), (1, False, False), (1, True, True)):
            ws = torch.zeros(C.lib.tn_convpool_bwd_workspace_bytes(B, *geom, need_dx) // 4 + 1, device='cuda')
            first, bad = None, {'dW': 0, 'db': 0, 'dx': 0}
            for r in range(reps):
                dW, db, dx = torch.full_like(W, 7.), torch.full_like(b, 7.), torch.full_like(x, 7.)
                C.call('tn_convpool_bwd', C.ptr(x), None if use_tie else C.ptr(a_full), C.ptr(tie) if use_tie else None,
                       C.ptr(pooled), C.ptr(dtop), C.ptr(W),
                       C.ptr(dW), C.ptr(db), C.ptr(dx) if need_dx else None, C.ptr(x) if below else None, C.ptr(ws),
                       B, Cin, S, M, 3, 0, O_, act, nn, 2, P, *(C.act_code('relu07') if below else (0, 0)), None)
                torch.cuda.synchronize()
                cur = (dW.clone(), db.clone(), dx.clone())
                if r == 0:
                    nans = [int(torch.isnan(t).sum()) for t in cur]
                    if any(nans):
                        print('   NaNs in (dW, db, dx):', nans, 'first dW nan at', torch.isnan(cur[0]).nonzero()[:4].tolist())
                if first is None:
                    first = cur
                else:
                    for name, u, v in zip(('dW', 'db', 'dx'), first, cur):
                        if not torch.equal(u, v):
                            bad[name] += 1
                            if bad[name] == 1:
                                d = (u != v).nonzero()
                                print('   first mismatch', name, 'rep', r, 'count', len(d), 'at', d[:3].tolist(),
                                      u[u != v][:3].tolist(), v[u != v][:3].tolist())
            print('case', (B, Cin, S, M), 'need_dx', need_dx, 'use_tie', use_tie, 'below', below, 'coop', os.environ.get('TN_SMALL_COOP', '1'),
                  'mismatching reps:', bad, flush=True)


if __name__ == '__main__':
    main()
