"""Data-parallel host logic on CPU: two gloo ranks, each computing its shard of the minibatch with
the oracle as the stand-in engine, reproduce the single-process gradient and update through
DistContext.all_reduce_sum (the same call the GPU path makes on its flat gradient buffer)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def flat(grads):
    return np.concatenate([g.ravel() for gg in grads if gg is not None for g in gg]).astype(np.float32)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    from oracle import theanet_oracle as O
    from theanet_b200.dist import init_from_env, shard_bounds
    import make_golden as MG
    ctx = init_from_env('cpu')
    assert (ctx.rank, ctx.world) == (rank, world)
    B = 8
    p = MG.load_prms('mnist.prms', B, 28)
    x, y = MG.synth(2 * B, 1, 28, 10, 1234)
    on = O.OracleNet(p['layers'], p['training_params'])
    out = []
    for step in range(2):
        lo, hi = shard_bounds(step, B, rank, world)
        # sample0 = first GLOBAL sample of the shard: masks are independent of the world size
        cost, lp = on.train_step(x[lo:hi], y[lo:hi], step=step, sample0=lo - step * B,
                                 global_batch=B, apply_update=False)
        buf = torch.from_numpy(np.concatenate([flat(on.last_grads), [np.float32(cost)]]))
        ctx.all_reduce_sum(buf)                        # the one collective of the step
        g = buf.numpy()
        k, grads = 0, []
        for gg in on.last_grads:                       # unflatten and apply the identical update
            if gg is None:
                grads.append(None)
                continue
            cur = []
            for t in gg:
                cur.append(g[k:k + t.size].reshape(t.shape))
                k += t.size
            grads.append(cur)
        on.apply_update(grads)
        out.append(float(g[-1]))
    q.put((rank, out, [t.copy() for L in on.spec for t in (L['params'] or [])]))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_data_parallel_matches_single_process():
    from oracle import theanet_oracle as O
    import make_golden as MG
    B = 8
    p = MG.load_prms('mnist.prms', B, 28)
    x, y = MG.synth(2 * B, 1, 28, 10, 1234)
    ref = O.OracleNet(copy.deepcopy(p['layers']), copy.deepcopy(p['training_params']))
    costs = [float(ref.train_step(x[s * B:(s + 1) * B], y[s * B:(s + 1) * B], step=s, sample0=0)[0])
             for s in range(2)]
    want = [t for L in ref.spec for t in (L['params'] or [])]

    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, c, wts in res:
        # the reported cost is the sum of the two shard costs (each already divided by the
        # global batch); weight-cost terms are zero in mnist.prms
        assert np.allclose(c, costs, rtol=1e-5)
        for a, b in zip(wts, want):
            assert np.max(np.abs(a - b)) <= 1e-5 * max(np.max(np.abs(b)), 1e-30)
    # both ranks hold bit-identical replicas
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)


def test_shard_bounds():
    from theanet_b200.dist import shard_bounds
    assert shard_bounds(0, 1024, 0, 8) == (0, 128)
    assert shard_bounds(3, 1024, 7, 8) == (3 * 1024 + 7 * 128, 4 * 1024)
    with pytest.raises(AssertionError):
        shard_bounds(0, 10, 0, 4)


def _facade_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    from theanet_b200.dist import init_from_env
    from theanet_b200.neuralnet import NeuralNet
    import make_golden as MG
    ctx = init_from_env('cpu')
    p = MG.load_prms('mnist.prms', 16, 28)
    p['training_params']['SEED'] = 100 + rank          # replicas must not depend on the local seed
    net = NeuralNet(p['layers'], p['training_params'], device='cpu', dist=ctx)
    net.step_count = 5
    net._set_ctl(48)
    from theanet_b200 import _C
    c = net._ctl_np
    q.put((rank, net.local_bsz, net.batch_sz, net.theta.numpy().copy(),
           (int(c[_C.CTL_STEP]), int(c[_C.CTL_SAMPLE0]), int(c[_C.CTL_ROW0])),
           float(np.float32(net.cur_learn_rate.get_value())), net._bucket_split(), int(net.n_flat)))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_facade_shards_the_batch_and_starts_from_rank0_weights():
    """NeuralNet under a 2-rank gloo group on the CPU (construction only -- there is no CPU
    execution path): BATCH_SZ stays the global minibatch, each rank owns BATCH_SZ / world samples
    starting at global sample rank * local (what keys the Philox masks), and every replica starts
    from rank 0's parameters whatever its local SEED drew."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_facade_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (r0, l0, b0, th0, ctl0, lr0, bk0, nf0), (r1, l1, b1, th1, ctl1, lr1, bk1, nf1) = res
    assert (l0, b0, l1, b1) == (8, 16, 8, 16)
    # gradient bucketing: the trailing dense layers (hidden, softmax) start at layer 5 / flat offset
    # 36 + 4 + 720 + 20 = 780; everything from there to the end (+ the NLL slot) is the early bucket
    assert bk0 == bk1 == (5, 780) and nf0 == 780 + 360000 + 500 + 5000 + 12
    assert np.array_equal(th0, th1) and np.abs(th0).max() > 0
    assert ctl0 == (5, 0, 48) and ctl1 == (5, 8, 48)
    assert lr0 == lr1
