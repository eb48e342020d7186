#!/usr/bin/env python
"""Benchmark of the theanet CNN-training hot path on B200 (BASELINE.json metric: images/sec of the
full training step on mnist.prms-shaped synthetic 28x28x1 batches, 1024 images per GPU, fp32).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      CPU reference arm (oracle port)

One step = elastic distortion + forward + softmax-NLL + backward + (N>1: one NCCL all-reduce of
the flat gradient buffer) + fused momentum/maxnorm update over one minibatch.  Prints ONE JSON
line (rank 0).  `value` is timed with the corpus resident in HBM; `e2e` goes through the public
get_trin_model() callable with HOST (pinned) buffers: per step one host->device copy of the
minibatch and a device->host read of cost + log-probabilities.
"""
import argparse
import ast
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 1024
IMG, NCLS = 28, 10
CORPUS_BATCHES = 64                # 64 x 1024 x 3136 B = 205 MB per GPU-share: larger than L2
WORKLOAD = "params/mnist.prms synthetic 28x28x1, batch 1024 per GPU, fp32 (BASELINE configs[1])"
# The other single-GPU configurations of BASELINE.json, reported next to the headline in
# `other_configs` (each with its own step time and step-level roofline); `nb` minibatches per GPU
# are cycled so that the corpus share exceeds the 126 MB L2.
CONFIGS = {
    'c2': dict(prms='mnist.prms', img=28, maps=1, ncls=10, per_gpu=1024, nb=64, bf16=False, dtype='f32',
               workload=WORKLOAD),
    'c3': dict(prms='3flat.prms', img=28, maps=1, ncls=457, per_gpu=1024, nb=64, bf16=False, dtype='f32',
               workload="params/3flat.prms (784 -> 1000 -> 457, no conv) synthetic 28x28x1, batch 1024 "
                        "per GPU, fp32 (BASELINE configs[2])"),
    'c4': dict(prms='cifar3conv.prms', img=32, maps=3, ncls=10, per_gpu=1024, nb=12, bf16=True, dtype='bf16',
               workload="CIFAR-shaped 32x32x3, 3-conv params/cifar3conv.prms, bf16 tensor-core conv "
                        "stack, batch 1024 per GPU (BASELINE configs[3])"),
    'c5': dict(prms='mnist.prms', img=64, maps=1, ncls=10, per_gpu=512, nb=20, bf16=False, dtype='f32',
               workload="ElasticLayer on, 64x64x1 synthetic, mnist.prms topology, batch 512 per GPU "
                        "(4096 at 8 GPUs), fp32 (BASELINE configs[4])"),
}
# algorithmic work of one training step, per image (layer-boundary convention)
FLOP_PER_IMG = 2810064
BYTES_PER_IMG = 181144
PARAM_BYTES_PER_STEP = 8 * 4 * 366290


def load_prms(global_batch, cfg='c2'):
    c = CONFIGS[cfg]
    with open(os.path.join(ROOT, 'params', c['prms'])) as f:
        p = ast.literal_eval(f.read())
    p['training_params'].update(SEED=555555, BATCH_SZ=global_batch)
    if c['bf16']:
        p['training_params']['CONV_DTYPE'] = 'bfloat16'
    p['layers'][0][1]['img_sz'] = c['img']
    return p


def synth_corpus(n, cfg='c2'):
    """uniform pixels thresholded so ~80% are exactly 0, uniform labels."""
    c = CONFIGS[cfg]
    rng = np.random.default_rng(1234)
    x = rng.random((n, c['maps'], c['img'], c['img']), dtype=np.float32)
    x *= (x > .8)
    y = rng.integers(0, c['ncls'], n).astype(np.int32)
    return x, y


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind each entry point,
    from profiles/r2_kernels.json -- written by tools/profile_kernels.py out of an `ncu --set full`
    capture of THIS workload; the file records the git revision of the binaries it profiled.
    Returns ({(entry point, ordinal): bytes}, provenance)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_kernels.json')) as f:
            d = json.load(f)
        tr = {(k['entry'], int(k['ordinal'])): k['dram_bytes'] for k in d['kernels'] if k.get('entry')}
        return tr, "profiles/r2_kernels.json (ncu --set full, git {})".format(d.get('git', '?'))
    except Exception:
        return {}, None


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            m = json.load(f)
        return float(m['hbm_gbs']), float(m['bf16_tflops']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits'], capture_output=True,
                                   text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == 'Active' for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_time(budget_s, batch, steps=None, warmup=1):
    """images/s of oracle.OracleNet.train_step (numpy + BLAS, all host threads) on `batch`-image
    minibatches of the same workload."""
    from oracle import theanet_oracle as O          # checker / baseline only
    p = load_prms(batch)
    on = O.OracleNet(p['layers'], p['training_params'])
    x, y = synth_corpus(batch * 2)
    for s in range(warmup):
        on.train_step(x[:batch], y[:batch], step=s)
    t0 = time.perf_counter()
    n = 0
    while True:
        i = n % 2
        on.train_step(x[i * batch:(i + 1) * batch], y[i * batch:(i + 1) * batch], step=warmup + n)
        n += 1
        if steps is not None:
            if n >= steps:
                break
        elif time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n * batch / dt, n, dt


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = PER_GPU_BATCH                         # same minibatch size as the GPU arm
    steps = max(1, min(args.steps, 24))           # bounded sample: a step takes ~0.15 s on 16 cores
    v, n, dt = cpu_step_time(None, batch, steps=steps, warmup=max(1, min(args.warmup, 3)))
    cores = host_threads()
    sample = "{} oracle train steps of {} images (numpy/BLAS im2col+SGEMM restatement)".format(n, batch)
    line = {
        "impl": "reference", "metric": "images/sec", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": batch, "per_gpu_batch": batch,
                   "cpu_sample_steps": n, "same_config": True},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# per-entry-point algorithmic work (for the dominant-kernel roofline)
# ------------------------------------------------------------------------------------------------
def stage_work(net):
    """(entry point, ordinal in call order) -> (algorithmic bytes, flops) of that call for the
    layers of `net` (every tensor a call must read or write, once)."""
    from theanet_b200.layer import ConvLayer, HiddenLayer, PoolLayer
    B = net.local_bsz
    L = net.tr_layers
    w, count = {}, {}

    def add(name, byts, fl):
        k = count.get(name, 0)
        count[name] = k + 1
        w[(name, k)] = (byts, fl)

    fwd, bwd = [], []
    for li, l in enumerate(L):
        if isinstance(l, ConvLayer):
            xin = B * l.num_prev_maps * l.in_sz ** 2 * 4
            a = B * l.num_maps * l.out_sz ** 2 * 4
            wt = l.W.size * 4
            fl = 2 * B * l.num_maps * l.out_sz ** 2 * l.num_prev_maps * l.filter_sz ** 2
            if li in getattr(net, 'conv_small', {}):
                # second-generation kernels: the un-pooled activations never reach HBM; the forward
                # writes the pooled map + one tie byte per pooled cell, the single backward launch
                # reads x, tie, pooled, dL/dpooled, W and writes dW, db (and dx)
                p = B * l.num_maps * net.conv_fused[li].out_sz ** 2 * 4
                fwd.append(('tn_convpool_fprop_train', xin + p + p // 4 + wt, fl))
                bwd.append((li, [('tn_convpool_bwd', xin + p // 4 + 2 * p + 2 * wt +
                                  (xin if net.need_below[li] else 0),
                                  fl * (2 if net.need_below[li] else 1))]))
            elif li in net.conv_fused:
                p = B * l.num_maps * net.conv_fused[li].out_sz ** 2 * 4
                fwd.append(('tn_convpool_fprop', xin + a + p + wt, fl))
                bwd.append((li, [('tn_convpool_bwd_weights', xin + a + 2 * p + wt, fl)] +
                            ([('tn_convpool_bwd_data', a + 2 * p + wt + xin, fl)]
                             if net.need_below[li] else [])))
            else:
                fwd.append(('tn_conv2d_fprop', xin + a + wt, fl))
                bwd.append((li, [('tn_conv2d_wgrad', xin + a + wt, fl)] +
                            ([('tn_conv2d_dgrad', xin + a + wt, fl)] if net.need_below[li] else [])))
        elif isinstance(l, PoolLayer) and (li - 1) not in net.conv_fused:
            xin = B * l.num_maps * l.in_sz ** 2 * 4
            yout = B * l.num_maps * l.out_sz ** 2 * 4
            fwd.append(('tn_maxpool_fwd', xin + yout, 0))
            bwd.append((li, [('tn_maxpool_bwd', 2 * xin + 2 * yout, 0)]))
        elif isinstance(l, HiddenLayer):
            a, b_, c = B * l.n_in * 4, l.n_in * l.n_out * 4, B * l.n_out * 4
            fl = 2 * B * l.n_in * l.n_out
            if net.head and li == len(L) - 1:
                fwd.append(('tn_softmax_head_fwd_bwd', 2 * a + b_ + 3 * c, 2 * fl))
                bwd.append((li, [('tn_softmax_head_bwd_weights', a + b_ + c, fl)]))
            else:
                fwd.append(('tn_dense_fwd', a + b_ + c, fl))
                bwd.append((li, [('tn_dense_bwd_weights', a + b_ + c, fl)] +
                            ([('tn_dense_bwd_data', 2 * a + b_ + c, fl)] if net.need_below[li] else [])))
    for name, byts, fl in fwd:
        add(name, byts, fl)
    for li, calls in reversed(bwd):
        for name, byts, fl in calls:
            add(name, byts, fl)
    return w


def profile_stages(net, fn, nb, reps):
    """Time every C-ABI call of the training step with CUDA events on the launching stream
    (eager execution, graphs off), `reps` steps.  Returns {(entry point, ordinal): mean ms}."""
    import torch
    from theanet_b200 import _C
    records = {}
    orig = _C.call
    counter = {}

    def timed_call(name, *a):
        key = name[:-3] if name.endswith('_sm') else name     # tn_dense_bwd_*_sm: same product, SM-bounded
        k = counter.get(key, 0)
        counter[key] = k + 1
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig(name, *a)
        e.record()
        records.setdefault((key, k), []).append((s, e))

    import theanet_b200.neuralnet as nnmod
    saved, saved_ov = net.use_graph, net.overlap_wgrad
    net.use_graph, net.overlap_wgrad = False, False
    _C.call = timed_call
    nnmod._C.call = timed_call
    try:
        for r in range(reps + 1):
            counter.clear()
            if r == 0:
                records.clear()
            fn(r % nb)
        torch.cuda.synchronize()
    finally:
        _C.call = orig
        nnmod._C.call = orig
        net.use_graph, net.overlap_wgrad = saved, saved_ov
    out = {}
    for k, evs in records.items():
        ts = [s.elapsed_time(e) for s, e in evs[1:]] or [evs[0][0].elapsed_time(evs[0][1])]
        out[k] = float(np.mean(ts))
    return out



def step_work(net):
    """(bytes, flops) of one training step in the layer-boundary convention of SURVEY.md 8(d): every
    layer reads its input and writes its output once in the forward pass, reads dL/dout (+ what its
    derivative needs) and writes dL/din in the backward pass; 32 bytes per parameter (two reads by
    the products, gradient write, theta/velocity/gradient read and theta/velocity write of the
    update); 2 FLOP per multiply-add, three products per weighted layer (two for the first)."""
    from theanet_b200.layer import ConvLayer, HiddenLayer
    B = net.local_bsz
    byts = fl = 0
    first = True
    for li, l in enumerate(net.tr_layers):
        o = net.out[li]
        if o is None or li == 0:
            if li == 0 and o is not None:
                byts += 2 * o.numel() * o.element_size()          # corpus read + warped image write
            continue
        n_in = net.out[li - 1].numel() * net.out[li - 1].element_size() if net.out[li - 1] is not None else 0
        n_out = o.numel() * o.element_size()
        if li in getattr(net, 'conv_tc', {}):                    # bf16 NHWC inside the tensor-core stack
            n_out //= 2
        byts += n_in + n_out                                       # forward
        byts += 2 * n_out + (n_in if net.need_below[li] else 0)    # backward
        macs = 0
        if isinstance(l, ConvLayer):
            macs = B * l.num_maps * l.out_sz ** 2 * l.num_prev_maps * l.filter_sz ** 2
        elif isinstance(l, HiddenLayer):
            macs = B * l.n_in * l.n_out
        if macs:
            fl += 2 * macs * (2 if first else 3)
            first = False
    byts += 32 * int(net.theta.numel())
    return byts, fl


def time_config(cfg, world, rank, dev, ctx, steps, warm, barrier, max_over_ranks):
    """One BASELINE configuration through the same timed loop as the headline: `warm` untimed steps,
    `steps` steps between CUDA events, max over ranks.  Returns the `other_configs` entry."""
    import torch
    from theanet_b200.neuralnet import NeuralNet
    c = CONFIGS[cfg]
    gb = c['per_gpu'] * world
    prms = load_prms(gb, cfg)
    x, y = synth_corpus(c['nb'] * gb, cfg)
    net = NeuralNet(prms['layers'], prms['training_params'], device=dev, dist=ctx)
    fn = net.get_trin_model(x, y, lazy=True)
    for s in range(warm):
        fn(s % c['nb'])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        fn((warm + s) % c['nb'])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    cost = float(net.cost.item())
    assert np.isfinite(cost), (cfg, cost)
    byts, fl = step_work(net)
    hbm, tfl, which = peaks()
    t_hbm, t_tc = byts / (hbm * 1e9), fl / (tfl * 1e12)
    bound = 'tensor' if (c['bf16'] and t_tc > t_hbm) else 'hbm'
    ach = (fl / (ms * 1e-3) / 1e12) if bound == 'tensor' else (byts / (ms * 1e-3) / 1e9)
    peak = tfl if bound == 'tensor' else hbm
    out = {"workload": c['workload'], "dtype": c['dtype'], "global_batch": gb, "per_gpu_batch": c['per_gpu'],
           "ms_per_step": ms, "images_per_s": gb / (ms * 1e-3), "steps": steps, "warmup": warm,
           "launches_per_step": int(net.launches.get('train', 0)), "cost_after": cost,
           "l2": "{} minibatches ({} MB per GPU) cycled".format(
               c['nb'], c['nb'] * c['per_gpu'] * c['maps'] * c['img'] ** 2 * 4 // 10 ** 6),
           "roofline": {"scope": "whole step, layer-boundary convention (SURVEY.md 8d)", "bound": bound,
                        "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == 'tensor' else "GB/s",
                        "frac": ach / peak, "algorithmic_bytes_per_step": byts,
                        "algorithmic_flops_per_step": fl,
                        "hbm_frac": byts / (ms * 1e-3) / 1e9 / hbm,
                        "tensor_frac": fl / (ms * 1e-3) / 1e12 / tfl, "peak_source": which}}
    del fn, net
    torch.cuda.empty_cache()
    return out


def dp_parity(world, rank, dev, ctx, steps=3):
    """Data-parallel correctness where the driver can see it: N ranks train `steps`
    steps of the headline workload from a fixed seed; rank 0 then trains the SAME global minibatches
    alone.  Reports the worst per-tensor deviation of the N-rank parameters from the single-rank
    ones (they differ only by the order of the gradient sum) and whether all replicas hold
    bit-identical parameters."""
    import torch
    import torch.distributed as dist
    from theanet_b200.neuralnet import NeuralNet, DistContext
    gb = PER_GPU_BATCH * world
    x, y = synth_corpus(steps * gb)

    def run(context):
        prms = load_prms(gb)
        net = NeuralNet(prms['layers'], prms['training_params'], device=dev, dist=context)
        fn = net.get_trin_model(x, y, lazy=True)
        for s in range(steps):
            fn(s)
        torch.cuda.synchronize(dev)
        return net

    net_n = run(ctx)
    theta_n = net_n.theta.detach().clone()
    digest = torch.stack([theta_n.double().sum(), theta_n.view(torch.int32).long().sum().double()])
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    identical = all(bool(torch.equal(all_d[0], d)) for d in all_d)
    res = None
    if rank == 0:
        net_1 = run(DistContext())
        worst = 0.0
        for seg_n, seg_1 in zip(net_n.get_init_params()['allwts'], net_1.get_init_params()['allwts']):
            for u, v in zip(seg_n, seg_1):
                d = float(np.max(np.abs(u.astype(np.float64) - v)) / max(float(np.max(np.abs(v))), 1e-30))
                worst = max(worst, d)
        res = {"steps": steps, "global_batch": gb, "max_rel": worst, "replicas_identical": identical,
               "tolerance": 1e-3,
               "what": "parameters after {} steps on {} ranks vs one rank on the same global "
                       "minibatches (max |a-b| / max |b| per tensor, worst tensor)".format(steps, world)}
        del net_1
    dist.barrier()
    del net_n
    torch.cuda.empty_cache()
    return res

# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from theanet_b200.neuralnet import NeuralNet, DistContext

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, "--gpus must match the torchrun world size"
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    ctx = DistContext()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        ctx = DistContext(rank, world, None)

    gb = PER_GPU_BATCH * world
    prms = load_prms(gb)
    x, y = synth_corpus(CORPUS_BATCHES * gb)
    net = NeuralNet(prms['layers'], prms['training_params'], device=dev, dist=ctx)
    fn = net.get_trin_model(x, y, lazy=True)
    nb = CORPUS_BATCHES
    warm = max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for s in range(warm):
        fn(s % nb)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = net.step_count
    e0.record()
    for s in range(args.steps):
        fn((warm + s) % nb)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    # K steps of this workload last tens of milliseconds, less than one nvidia-smi query: keep the
    # identical loop running (untimed) until the sampler has a handful of readings under load
    for s_tail in range(max(0, 3000 - args.steps)):      # same count on every rank (collectives)
        fn((warm + args.steps + s_tail) % nb)
    torch.cuda.synchronize(dev)
    clocks = sampler.summary() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = "the timed region and an untimed continuation of the same loop"
    launches = int(net.launches.get('train', 0)) * args.steps
    value = args.steps * gb / (ms * 1e-3)

    # ---- end to end through the public API with host buffers --------------------------------
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.from_numpy(y).pin_memory()
    fn_e2e = net.get_trin_model(xh, yh, resident=False, lazy=False)
    for s in range(warm):
        fn_e2e(s % nb)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        cost, feats, _ = fn_e2e((warm + s) % nb)
    torch.cuda.synchronize(dev)
    dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    Bl = net.local_bsz
    e2e = {"value": args.steps * gb / dt, "unit": "images/s",
           "h2d_bytes_per_step": world * (Bl * IMG * IMG * 4 + Bl * 4 + 32),
           "d2h_bytes_per_step": world * (4 + Bl * NCLS * 4),
           "ms_per_step": 1e3 * dt / args.steps,
           "api": "NeuralNet.get_trin_model(x_host, y_host, resident=False)(batch_index)"}
    assert np.isfinite(float(cost))

    # ---- dominant kernel, timed live with CUDA events ------------------------------------------
    roof = None
    hbm, tfl, which = peaks()
    if rank == 0 or world > 1:
        fn_prof = net.get_trin_model(x, y, lazy=True)
        stages = profile_stages(net, fn_prof, nb, reps=10)
        if rank == 0:
            work = stage_work(net)
            traffic, traffic_src = ncu_traffic()
            total_ms = sum(stages.values())
            (name, k), t_ms = max(stages.items(), key=lambda kv: kv[1])
            li = k
            byts, fl = work.get((name, k), (None, None))
            share = {"{}#{}".format(n_, k_): round(v / total_ms, 4)
                     for (n_, k_), v in sorted(stages.items(), key=lambda kv: -kv[1])[:8]}
            if byts is not None:
                ach = byts / (t_ms * 1e-3) / 1e9
                roof = {"bound": "hbm", "kernel": "{} (call #{} of the step)".format(name, li),
                        "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                        "traffic": traffic.get((name, k)) if world == 1 else None,
                        "traffic_source": traffic_src,
                        "launch_ms": t_ms, "algorithmic_bytes": byts,
                        "algorithmic_flops": fl,
                        "achieved_tflops": (fl / (t_ms * 1e-3) / 1e12) if fl else None,
                        "peak_source": which, "step_share": share,
                        "eager_step_ms_sum": total_ms}
            else:
                roof = {"bound": "hbm", "kernel": name, "achieved": None, "peak": hbm,
                        "unit": "GB/s", "frac": None, "traffic": None, "launch_ms": t_ms,
                        "peak_source": which, "step_share": share}
    barrier()

    # ---- the other BASELINE configurations, and data-parallel parity at N > 1 ---------------------
    del fn, fn_e2e
    torch.cuda.empty_cache()
    others = {}
    if not args.no_other:
        k_other = max(10, min(args.steps, 50))
        for cfg in (('c3', 'c4', 'c5') if world == 1 else ('c4', 'c5')):
            try:
                others[cfg] = time_config(cfg, world, rank, dev, ctx, k_other, warm, barrier,
                                          max_over_ranks)
            except Exception as ex:                       # a config must not cost the headline
                others[cfg] = {"error": "{}: {}".format(type(ex).__name__, ex)}
                barrier()
    parity = dp_parity(world, rank, dev, ctx) if world > 1 else None

    if rank != 0:
        finish(world, dev)
        return
    # whole-step view against the HBM roofline (SURVEY.md 8d figures)
    step_bytes = BYTES_PER_IMG * PER_GPU_BATCH + PARAM_BYTES_PER_STEP
    step_gbs = step_bytes / (ms / args.steps * 1e-3) / 1e9

    cpu_v, cpu_n, cpu_dt = (None, 0, 0.0)
    cpu = None
    if world == 1:
        cpu_v, cpu_n, cpu_dt = cpu_step_time(12.0, PER_GPU_BATCH)
        cpu = {"value": cpu_v, "unit": "images/s", "cores": host_threads(), "kind": "port",
               "sample": "{} oracle train steps of {} images in {:.1f} s (numpy/BLAS "
                         "restatement of the Theano CPU path, same minibatch size as the GPU arm)".format(
                             cpu_n, PER_GPU_BATCH, cpu_dt)}

    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": gb, "per_gpu_batch": PER_GPU_BATCH,
                   "parallelism": "dp{}".format(world),
                   "l2": "corpus of {} minibatches ({} MB per GPU) cycled, larger than the 126 MB "
                         "L2; no explicit flush".format(nb, nb * PER_GPU_BATCH * IMG * IMG * 4 // 10 ** 6),
                   "cuda_graph": bool(net.use_graph),
                   "collective": ("none" if world == 1 else
                                  "hybrid, inside the step's CUDA graph: NCCL all-reduce of the dense-layer "
                                  "gradients overlapped with the conv backward pass + the late conv "
                                  "gradients summed over CUDA-IPC peer memory (NVLink) in the update kernel"
                                  if getattr(net, 'dp_hybrid', False) else
                                  "fused into the update kernel over CUDA-IPC peer memory (NVLink)"
                                  if net.dp_fused else
                                  "NCCL all-reduce of the flat gradient buffer"
                                  + (" (in the CUDA graph, dense bucket overlapped)" if net.nccl_in_graph
                                     else " (eager, between two CUDA graphs)"))},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roof, "cpu_baseline": cpu, "other_configs": others, "dp_parity": parity,
        "step_roofline": {"bound": "hbm", "algorithmic_bytes_per_step": step_bytes,
                          "achieved": step_gbs, "peak": hbm, "unit": "GB/s",
                          "frac": step_gbs / hbm, "flop_per_img": FLOP_PER_IMG},
    }
    print(json.dumps(line), flush=True)
    finish(world, dev)


def finish(world, dev):
    """Leave without tearing the NCCL communicator down: CUDA graphs that captured collectives are
    still alive, and destroy_process_group() behind them has been seen to block."""
    import torch
    if world > 1:
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-other', action='store_true', help="skip other_configs (C3/C4/C5)")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
