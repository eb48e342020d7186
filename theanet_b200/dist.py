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
