"""GPU check of the tcgen05 dense path against float64 numpy: errors and CUDA-event timings of
tn_dense_fwd / tn_dense_bwd_data / tn_dense_bwd_weights per dense mode (1 = CUDA cores,
2 = tensor cores single TF32 pass, 3 = tensor cores 3xTF32 (cluster split-K), 4 = first-generation
3xTF32 kernel), plus a cuBLAS yardstick through torch.matmul (this tool only).   python tools/gemm_tc_check.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from theanet_b200 import _C as C  # noqa: E402

SHAPES = [(1024, 720, 500), (128, 784, 1000), (64, 100, 36), (20, 720, 500), (1024, 4500, 500),
          (1024, 1024, 1024)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]]


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def timeit(fn, reps=20):
    """Device time per call: `reps` calls captured into one CUDA graph (no host overhead)."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def st():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def main():
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(7)
    for (B, n_in, n_out) in SHAPES:
        x = rng.standard_normal((B, n_in)).astype(np.float32)
        W = (rng.standard_normal((n_in, n_out)) / np.sqrt(n_in)).astype(np.float32)
        b = rng.standard_normal(n_out).astype(np.float32)
        g = rng.standard_normal((B, n_out)).astype(np.float32)
        xd, Wd, bd, gd = (torch.from_numpy(a).to(dev) for a in (x, W, b, g))
        out = torch.zeros(B, n_out, device=dev)
        dx = torch.zeros(B, n_in, device=dev)
        dW = torch.zeros(n_in, n_out, device=dev)
        db = torch.zeros(n_out, device=dev)
        ctl = torch.zeros(8, dtype=torch.int32, device=dev)
        want_f = x.astype(np.float64) @ W.astype(np.float64) + b
        want_dx = g.astype(np.float64) @ W.astype(np.float64).T
        want_dW = x.astype(np.float64).T @ g.astype(np.float64)
        for mode in (1, 2, 4, 3):     # 3 = default (cluster split-K 3xTF32), 4 = first generation
            C.call('tn_set_dense_mode', mode)

            def f_fwd():
                C.call('tn_dense_fwd', C.ptr(xd), C.ptr(Wd), C.ptr(bd), C.ptr(out), B, n_in, n_out,
                       C.ACT_LINEAR, 0, 1.0, 0, C.ptr(ctl), None, 1.0, st())

            def f_dx():
                C.call('tn_dense_bwd_data', C.ptr(gd), C.ptr(Wd), C.ptr(dx), B, n_in, n_out, None,
                       0, 0, 1.0, 0, C.ptr(ctl), None, st())

            def f_dw():
                C.call('tn_dense_bwd_weights', C.ptr(xd), C.ptr(gd), C.ptr(dW), C.ptr(db), B, n_in,
                       n_out, st())

            res = {'shape': [B, n_in, n_out], 'mode': mode}
            for name, fn, t, want in (('fwd', f_fwd, out, want_f), ('dx', f_dx, dx, want_dx),
                                      ('dW', f_dw, dW, want_dW)):
                t.fill_(-77.0)
                fn()
                torch.cuda.synchronize()
                res[name + '_err'] = rel(t.cpu().numpy().astype(np.float64), want)
                res[name + '_us'] = round(timeit(fn), 2)
            print(json.dumps(res), flush=True)
        C.call('tn_set_dense_mode', 0)
        # yardstick (this tool only, never the product): cuBLAS through torch.matmul
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            res = {'shape': [B, n_in, n_out], 'mode': 'cuBLAS ' + ('tf32' if tf32 else 'fp32')}
            for name, fn, want in (('fwd', lambda: torch.matmul(xd, Wd, out=out), want_f - b),
                                   ('dx', lambda: torch.matmul(gd, Wd.t(), out=dx), want_dx),
                                   ('dW', lambda: torch.matmul(xd.t(), gd, out=dW), want_dW)):
                r = fn()
                torch.cuda.synchronize()
                res[name + '_err'] = rel(r.cpu().numpy().astype(np.float64), want)
                res[name + '_us'] = round(timeit(fn), 2)
            print(json.dumps(res), flush=True)
        torch.backends.cuda.matmul.allow_tf32 = False


if __name__ == '__main__':
    main()
