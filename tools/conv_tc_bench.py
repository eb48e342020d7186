"""Throughput of the tcgen05 implicit-GEMM convolution kernels at the C4 shapes (CUDA-graph timed):
TFLOP/s and fraction of the measured bf16 peak.   python tools/conv_tc_bench.py [B]"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from theanet_b200 import _C as C  # noqa: E402


def st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=10):
    if os.environ.get('TN_BENCH_PLAIN'):          # for ncu: two plain launches, no graph
        fn()
        fn()
        torch.cuda.synchronize()
        return 1.0
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
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops']
    except Exception:
        peak = 1590.0
    for (Cin, S, M) in [(64, 16, 128), (128, 8, 256), (256, 8, 256), (128, 16, 256)]:
        f, pad = 3, 1
        x = torch.randn((B, S, S, Cin), device='cuda').to(torch.bfloat16)
        gz = torch.randn((B, S, S, M), device='cuda').to(torch.bfloat16)
        W = torch.randn((M, Cin, f, f), device='cuda') / (Cin * 9) ** .5
        b = torch.zeros(M, device='cuda')
        Wp = torch.zeros(M * 9 * Cin, dtype=torch.bfloat16, device='cuda')
        Wpd = torch.zeros_like(Wp)
        C.call('tn_conv2d_tc_pack_weights', C.ptr(W), C.ptr(Wp), M, Cin, f, 0, st())
        C.call('tn_conv2d_tc_pack_weights', C.ptr(W), C.ptr(Wpd), M, Cin, f, 1, st())
        a = torch.zeros((B, S, S, M), dtype=torch.bfloat16, device='cuda')
        p = torch.zeros((B, S // 2, S // 2, M), dtype=torch.bfloat16, device='cuda')
        dx = torch.zeros((B, S, S, Cin), dtype=torch.bfloat16, device='cuda')
        dW, db = torch.zeros_like(W), torch.zeros_like(b)
        nb = C.lib.tn_conv2d_tc_wgrad_workspace_bytes(B, Cin, M, f, S)
        ws = torch.zeros(nb // 4 + 4, device='cuda')
        flop = 2.0 * B * S * S * M * Cin * 9
        res = {'shape': 'B%d C%d %dx%d M%d f3' % (B, Cin, S, S, M), 'gflop': round(flop / 1e9, 2)}
        for name, fn in (
                ('fprop', lambda: C.call('tn_conv2d_tc_fprop', C.ptr(x), C.ptr(Wp), C.ptr(b), C.ptr(a),
                                         C.ptr(p), B, Cin, S, M, f, pad, S, C.ACT_LEAKY, 5, st())),
                ('dgrad', lambda: C.call('tn_conv2d_tc_dgrad', C.ptr(gz), C.ptr(Wpd), C.ptr(dx), B, Cin, S,
                                         M, f, pad, S, st())),
                ('wgrad', lambda: C.call('tn_conv2d_tc_wgrad', C.ptr(x), C.ptr(gz), C.ptr(dW), C.ptr(db),
                                         C.ptr(ws), B, Cin, S, M, f, pad, S, st()))):
            t = timeit(fn)
            res[name + '_us'] = round(t * 1e6, 1)
            res[name + '_tflops'] = round(flop / t / 1e12, 1)
            res[name + '_frac_of_peak'] = round(flop / t / 1e12 / peak, 3)
        print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
