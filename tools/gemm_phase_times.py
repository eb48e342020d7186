"""Phase times inside the cluster split-K dense kernel (gemm_tc_sk_kernel): every CTA stamps
clock64() at its phase boundaries (tn_dense_debug_timestamps).   python tools/gemm_phase_times.py [BxKxN]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from theanet_b200 import _C as C  # noqa: E402

PH = ['setup (barriers, TMEM alloc)', 'first tile landed', 'split of last k-block done', 'last MMA done',
      'TMEM drain -> smem', 'cluster barrier', 'DSMEM reduce + epilogue']


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else '1024x720x500'
    B, n_in, n_out = (int(v) for v in shape.split('x'))
    dev = 'cuda'
    x = torch.randn(B, n_in, device=dev)
    W = torch.randn(n_in, n_out, device=dev)
    b = torch.randn(n_out, device=dev)
    g = torch.randn(B, n_out, device=dev)
    out, dx, dW, db = torch.zeros(B, n_out, device=dev), torch.zeros(B, n_in, device=dev), \
        torch.zeros(n_in, n_out, device=dev), torch.zeros(n_out, device=dev)
    ctl = torch.zeros(8, dtype=torch.int32, device=dev)
    buf = torch.zeros(1024 * 16, dtype=torch.int64, device=dev)
    calls = {
        'fwd': lambda: C.call('tn_dense_fwd', C.ptr(x), C.ptr(W), C.ptr(b), C.ptr(out), B, n_in, n_out,
                              C.ACT_LINEAR, 0, 1.0, 0, C.ptr(ctl), None, 1.0, None),
        'dx': lambda: C.call('tn_dense_bwd_data', C.ptr(g), C.ptr(W), C.ptr(dx), B, n_in, n_out, None, 0, 0,
                             1.0, 0, C.ptr(ctl), None, None),
        'dW': lambda: C.call('tn_dense_bwd_weights', C.ptr(x), C.ptr(g), C.ptr(dW), C.ptr(db), B, n_in,
                             n_out, None),
    }
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        buf.zero_()
        C.call('tn_dense_debug_timestamps', C.ptr(buf))
        fn()
        torch.cuda.synchronize()
        C.call('tn_dense_debug_timestamps', None)
        t = buf.cpu().numpy().reshape(-1, 16)
        t = t[t[:, 0] != 0]
        print('{}: {} CTAs'.format(name, len(t)))
        for i, label in enumerate(PH):
            d = (t[:, i + 1] - t[:, i]).astype(np.float64)
            print('   {:34s} cycles: mean {:8.0f}  max {:8.0f}'.format(label, d.mean(), d.max()))
        d = (t[:, 7] - t[:, 0]).astype(np.float64)
        print('   {:34s} cycles: mean {:8.0f}  max {:8.0f}'.format('whole CTA', d.mean(), d.max()))


if __name__ == '__main__':
    main()
