"""Scratch diagnosis (GPU): the dense products of the C5 network's first step, on the network's own
operands, per dense mode, against float64."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from theanet_b200 import _C as C                # noqa: E402
from theanet_b200.neuralnet import NeuralNet    # noqa: E402
import test_gpu_net as T                        # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def st():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def products(x, g, W, tag):
    dev = 'cuda'
    B, n_in = x.shape
    n_out = g.shape[1]
    xd, gd, Wd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, g, W))
    ctl = torch.zeros(8, dtype=torch.int32, device=dev)
    x64, g64, W64 = x.astype(np.float64), g.astype(np.float64), W.astype(np.float64)
    want_dx, want_dW = g64 @ W64.T, x64.T @ g64
    cond_dW = float(np.max(np.abs(x64).T @ np.abs(g64)) / np.max(np.abs(want_dW)))
    cond_dx = float(np.max(np.abs(g64) @ np.abs(W64).T) / np.max(np.abs(want_dx)))
    print(tag, 'cond dW {:.1f} dx {:.1f}; |x| max {:.3g} |g| max {:.3g} min nz {:.3g} zeros x {:.2f} g {:.2f}'.format(
        cond_dW, cond_dx, np.abs(x).max(), np.abs(g).max(), np.abs(g[g != 0]).min(),
        float((x == 0).mean()), float((g == 0).mean())))
    for mode in (1, 2, 3):
        C.call('tn_set_dense_mode', mode)
        dx = torch.zeros(B, n_in, device=dev)
        dW = torch.zeros(n_in, n_out, device=dev)
        db = torch.zeros(n_out, device=dev)
        C.call('tn_dense_bwd_data', C.ptr(gd), C.ptr(Wd), C.ptr(dx), B, n_in, n_out, None, 0, 0, 1.0, 0,
               C.ptr(ctl), None, st())
        C.call('tn_dense_bwd_weights', C.ptr(xd), C.ptr(gd), C.ptr(dW), C.ptr(db), B, n_in, n_out, st())
        torch.cuda.synchronize()
        print('   mode', mode, 'dx {:.2e} dW {:.2e}'.format(rel(dx.cpu().numpy(), want_dx),
                                                         rel(dW.cpu().numpy(), want_dW)))
    C.call('tn_set_dense_mode', 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    print('   cuBLAS fp32 dx {:.2e} dW {:.2e}'.format(rel((gd @ Wd.t()).cpu().numpy(), want_dx),
                                                     rel((xd.t() @ gd).cpu().numpy(), want_dW)))


def main():
    prms = T.load_prms('mnist.prms', 512, 64)
    x, y = T.synth(1024, 1, 64, 10)
    net = NeuralNet(prms['layers'], prms['training_params'], use_graph=False)
    fn = net.get_trin_model(x, y)
    fn(0)
    torch.cuda.synchronize()
    li = 5
    xin = net.out[li - 1].reshape(512, -1).cpu().numpy()
    g = net.dbuf[li].cpu().numpy() if net.dbuf[li] is not None else None
    W = net.tr_layers[li].w.tensor.cpu().numpy()
    print('x', xin.shape, 'g', None if g is None else g.shape, 'W', W.shape)
    products(xin, g.reshape(512, -1), W, 'network operands')
    rng = np.random.default_rng(0)
    products(xin, rng.standard_normal(g.reshape(512, -1).shape).astype(np.float32), W, 'x net, g normal')
    products(rng.standard_normal(xin.shape).astype(np.float32), g.reshape(512, -1), W, 'x normal, g net')
    gs = g.reshape(512, -1)
    products(xin, (gs * 2.0 ** 20).astype(np.float32), W, 'g net scaled 2^20')
    gn = rng.standard_normal(gs.shape).astype(np.float32)
    products(xin, (gn * 2.0 ** -20).astype(np.float32), W, 'g normal scaled 2^-20')
    products(xin, (gn * (gs != 0)).astype(np.float32), W, 'g normal with net zeros')


if __name__ == '__main__':
    main()
