"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPUs,
gloo for the CPU tests) -- the reference has no multi-device path at all.

The minibatch is split into contiguous per-rank shards; every rank computes the gradient of ITS
shard of the global-batch mean loss (divisor = global batch), one sum-all-reduce of the flat
gradient buffer reproduces the single-process gradient, and every rank applies the identical
update.  Random streams are keyed by the GLOBAL sample index, so masks do not depend on the world
size."""
import os

import torch
import torch.distributed as dist


class DistContext:
    def __init__(self, rank=0, world=1, group=None):
        self.rank, self.world, self.group = rank, world, group

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def broadcast(self, t, src=0):
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)


def shard_bounds(batch_index, global_batch, rank, world):
    """Rows [lo, hi) of the corpus that `rank` processes for minibatch `batch_index`."""
    assert global_batch % world == 0, "BATCH_SZ must divide by the world size"
    per = global_batch // world
    lo = batch_index * global_batch + rank * per
    return lo, lo + per


def init_from_env(device=None):
    """DistContext from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        return DistContext()
    rank = int(os.environ['RANK'])
    if not dist.is_initialized():
        if device is not None and torch.device(device).type == 'cuda':
            dist.init_process_group('nccl', device_id=torch.device(device))
        else:
            dist.init_process_group('gloo')
    return DistContext(rank, world, None)


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so that torch can view library-owned memory."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (ptr, False),
                                         'version': 2}


class PeerBuffers:
    """The double-buffered flat gradient buffers and the flag array of every rank, mapped into
    this process with CUDA IPC, for the fused all-reduce + update kernel
    (tn_allreduce_sgd_update).  Memory comes from tn_peer_alloc (plain cudaMalloc): an IPC handle
    must describe exactly one allocation, which a caching allocator does not guarantee."""

    def __init__(self, ctx, nfloats, device):
        import ctypes
        from . import _C
        self.ctx, self._C, self._opened, self._owned = ctx, _C, [], []
        W, me = ctx.world, ctx.rank

        def alloc(nbytes):
            p = ctypes.c_void_p()
            _C.call('tn_peer_alloc', nbytes, ctypes.byref(p))
            self._owned.append(p.value)
            return p.value

        def handle(ptr):
            buf = ctypes.create_string_buffer(64)
            _C.call('tn_ipc_get_handle', ptr, buf)
            return buf.raw

        own = [alloc(4 * nfloats), alloc(4 * nfloats), alloc(4 * 32)]   # flags: update kernel words 0..8, all-reduce kernel words 16..25
        self.grad = [torch.as_tensor(_RawCuda(own[k], nfloats, '<f4'), device=device) for k in (0, 1)]
        self.flags = torch.as_tensor(_RawCuda(own[2], 32, '<i4'), device=device)
        torch.cuda.synchronize(device)
        mine = [handle(p) for p in own]
        everyone = [None] * W
        dist.all_gather_object(everyone, mine, group=ctx.group)
        ptrs = []
        for r in range(W):
            if r == me:
                ptrs.append(own)
                continue
            row = []
            for h in everyone[r]:
                p = ctypes.c_void_p()
                _C.call('tn_ipc_open_handle', ctypes.create_string_buffer(h, 64), ctypes.byref(p))
                self._opened.append(p.value)
                row.append(p.value)
            ptrs.append(row)
        arr = ctypes.c_void_p * W
        self.grad_ptrs = [arr(*[ptrs[r][k] for r in range(W)]) for k in (0, 1)]   # per parity
        self.flag_ptrs = arr(*[ptrs[r][2] for r in range(W)])
        dist.barrier(group=ctx.group)

    def close(self):
        for p in self._opened:
            self._C.lib.tn_ipc_close_handle(p)
        self._opened = []
