"""Data parallelism for the SAVP path: one process per GPU, NCCL all-reduce (NVLink 5 / NVSwitch) of the two flat
gradient buffers -- discriminator after the D backward, generator after the G backward (the reference issues one
tf.contrib.nccl.all_sum per variable, tf_utils.py:450-480, base_model.py:590-592, 614-616).  The 1/world of the mean
(tf_utils.py:473-474) is folded into the Adam kernel.  Every op of the model is per-sample (instance norm, CDNA
kernels, clip sampling), so no activation ever crosses GPUs."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torch.distributed.run).  Returns (rank, local_rank, world)."""
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kwargs['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kwargs)
    return rank, local, world


def shard_batch(global_batch, rank, world):
    """Reference semantics (base_model.py:523-527): the global batch is split into `world` equal shards."""
    if global_batch % world:
        raise ValueError('batch size %d is not divisible by the number of GPUs %d' % (global_batch, world))
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


def broadcast_state(flat_buffers, src=0):
    """Replicas start from identical variables (base_model.py:640-646 copies tower 0's values)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for buf in flat_buffers:
            dist.broadcast(buf, src)


def make_allreduce():
    """Returns a callable summing a flat gradient buffer over ranks in place (None when world == 1)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None

    def allreduce(buf):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return allreduce


def mean_scalars(t):
    """Logged scalar losses are averaged over replicas (tf_utils.py:489-490)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= dist.get_world_size()
    return t


class GradientReducer(object):
    """Bucketed, overlapped gradient all-reduce (sum) of a flat buffer.  `reduce_range(lo, hi)` enqueues the all-reduce of
    flat[lo:hi] on ONE dedicated communication stream as soon as the producing stream has finished that range (an event), so
    the exchange of one layer's gradients runs under the computation of the next; NCCL operations of a communicator must
    not run concurrently, which the single stream guarantees (also inside a captured CUDA graph, where they become a chain of
    nodes).  `join()` makes the current stream wait for everything enqueued.  The reference issues one nccl.all_sum per
    variable after the whole backward pass (tf_utils.py:450-480); the sums are identical."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)

    def reduce_range(self, flat, lo, hi):
        if hi <= lo:
            return
        ev = torch.cuda.Event()
        ev.record()                                  # on the producing (current) stream
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM)

    def join(self):
        torch.cuda.current_stream().wait_stream(self.stream)
