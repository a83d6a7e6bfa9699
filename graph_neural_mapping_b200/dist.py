"""Data-parallel plumbing: one process per GPU over torch.distributed (NCCL on the GPU box,
gloo in CPU tests). The reference has no distributed code at all; this is new.

Graphs shard by contiguous slices of the driver's batch order. The only
exchanges on the data path are
  * BatchNorm batch statistics: one all-reduce of [sum, sumsq] per BatchNorm forward and of
    [sum dy, sum dy*xhat] per BatchNorm backward (sync-BN over the global batch),
  * the DGI negative rows: rank 0 owns global rows [0, B_global) of n_f; they are broadcast
    forward and their gradient is reduced back to rank 0,
  * the parameter gradients (averaged) after backward.
"""
import os

import torch
import torch.distributed as td


class Comm(object):
    """Communicator used by the engine; world == 1 turns every call into a no-op."""

    def __init__(self, group=None, world=1, rank=0):
        self.group = group
        self.world = int(world)
        self.rank = int(rank)

    def all_reduce_sum(self, t):
        if self.world > 1:
            td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)
        return t

    def broadcast(self, t, src):
        if self.world > 1:
            td.broadcast(t, src=src, group=self.group)
        return t

    def reduce_sum(self, t, dst):
        if self.world > 1:
            td.reduce(t, dst=dst, op=td.ReduceOp.SUM, group=self.group)
        return t


SINGLE = Comm()


def from_env():
    """Comm over the default process group if torch.distributed is initialised, else single."""
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        return Comm(None, td.get_world_size(), td.get_rank())
    return SINGLE


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK, WORLD_SIZE,
    LOCAL_RANK, MASTER_ADDR, MASTER_PORT). Returns (comm, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world <= 1:
        return SINGLE, local_rank
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        td.init_process_group(backend=backend, **kw)
    return Comm(None, td.get_world_size(), td.get_rank()), local_rank


def shard(batch_graph, comm):
    """This rank's contiguous slice of the driver's batch (order defines the DGI negatives)."""
    if comm.world == 1:
        return batch_graph
    b = len(batch_graph)
    if b % comm.world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (b, comm.world))
    per = b // comm.world
    return batch_graph[comm.rank * per:(comm.rank + 1) * per]


def average_gradients(model, comm):
    """All-reduce (mean) of every parameter gradient in one flat buffer. With each rank's loss
    being the mean over its own shard, the average equals the gradient of the global-batch mean
    loss of the single-process reference (main.py:34-41)."""
    if comm.world == 1:
        return
    ps = [p for p in model.parameters() if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    comm.all_reduce_sum(flat)
    flat.div_(comm.world)
    off = 0
    for p in ps:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def shutdown():
    """Tear the default process group down. Call `model.release_graphs()` first: CUDA graphs that captured NCCL
    kernels keep the communicator alive and `destroy_process_group` would wait for them."""
    import gc
    import sys
    import threading
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if td.is_available() and td.is_initialized():
        # belt and braces: results are already printed when this runs; never let a stuck communicator
        # teardown hold the job (and the GPUs) hostage
        sys.stdout.flush()
        sys.stderr.flush()
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        td.destroy_process_group()
        watchdog.cancel()
