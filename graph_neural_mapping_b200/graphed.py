"""CUDA-graph execution of the training step's encoder forward / backward.

A training step launches ~130 libgnm kernels plus a few dozen small torch ops; issued one by
one from Python they leave the B200 idle between launches (host-bound: 14-30 ms per step for
11 ms of kernels in the first measurements). The shapes of a training step are static
(B graphs of N nodes), so `engine.run_forward` / `engine.run_backward` - the same Python
orchestration, unchanged - are captured ONCE per shape into two CUDA graphs and replayed:

  per step (host):  3 small H2D copies (slot addresses, node offsets, DGI permutation)
                    -> forward graph replay -> [driver: heads, loss, autograd]
                    -> 2 D2D copies of the incoming gradients -> backward graph replay

Inputs and outputs live in static buffers owned by the plan; outputs and gradients are
cloned before they are handed to autograd so the caller never aliases graph memory.
BatchNorm running statistics, NCCL all-reduces (data parallel) and the batch assembly kernel
are part of the captured work.
"""
import torch

from . import engine as _engine
from . import ops as _ops


import contextlib
import gc


@contextlib.contextmanager
def quiet_capture():
    """Around a CUDA-graph capture: Python's cyclic garbage collector must not run inside it. Collecting an old plan
    (a CUDAGraph with its private memory pool, pinned staging buffers) calls cudaGraphExecDestroy / cudaFree-class APIs,
    which are not permitted while a stream is capturing and invalidate the capture ("operation not permitted when
    stream is capturing" at capture_end, seen once in a long test process). Collect first, then hold the collector."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


class StepPlan(object):
    """Static buffers + captured graphs for one (B, N, flags) training-step shape."""

    def __init__(self, model, host_batch, n_neg, comm):
        dev = model.eps.device
        h = host_batch
        self.model, self.comm, self.dev = model, comm, dev
        self.b, self.m = h.b, h.m
        self.nnz_cap = int(h.nnz * 1.02) + 4096
        self.packed = torch.zeros(5 * h.b + 1, dtype=torch.int64, device=dev)
        self.node_off = torch.zeros(h.b + 1, dtype=torch.int32, device=dev)
        self.neg_idx = torch.zeros(n_neg, dtype=torch.int32, device=dev)
        self.pool = torch.cuda.graph_pool_handle()
        self.fwd_graph = None
        self.bwd_graph = None
        self.generation = 0
        self.h = None
        self._stage = [None, None, None]
        self._stage_i = 0
        self.sig = self._signature(h)
        self.ptrs = self._pointers()

    # a captured graph bakes in kernel choices and pointers: anything that changes them is part of the key
    @staticmethod
    def _signature(h):
        return (h.b, h.m, h.uniform_n, h.onehot, h.feat_dim, h.n_max, h.dense, h.has_isolated, h.same_tags)

    def _pointers(self):
        ps = [p.data_ptr() for p in _engine.flat_params(self.model)]
        ps += [b.data_ptr() for b in _engine.flat_buffers(self.model)]
        return ps

    def compatible(self, h, n_neg):
        return (self._signature(h) == self.sig and h.nnz <= self.nnz_cap and n_neg == self.neg_idx.shape[0]
                and self._pointers() == self.ptrs)

    def load(self, h, perm):
        """Per-step host -> device traffic: slot addresses / offsets and the DGI permutation, staged through a small
        ring of pinned buffers and copied asynchronously, so the host never waits for the previous step's kernels and
        can prepare (and enqueue) the next step while the GPU is still busy with this one."""
        self.h = h
        i = self._stage_i
        self._stage_i = (i + 1) % len(self._stage)
        if self._stage[i] is None:
            self._stage[i] = (torch.empty(self.packed.shape[0], dtype=torch.int64, pin_memory=True),
                              torch.empty(self.node_off.shape[0], dtype=torch.int32, pin_memory=True),
                              torch.empty(self.neg_idx.shape[0], dtype=torch.int32, pin_memory=True),
                              torch.cuda.Event())
        else:
            self._stage[i][3].synchronize()           # the copies that last read this slot (three steps ago) are done
        s_packed, s_off, s_neg, ev = self._stage[i]
        s_packed.numpy()[:] = h.packed
        s_off.numpy()[:] = h.node_off
        s_neg.numpy()[:] = perm
        self.packed.copy_(s_packed, non_blocking=True)
        self.node_off.copy_(s_off, non_blocking=True)
        self.neg_idx.copy_(s_neg, non_blocking=True)
        ev.record()
        return h.packed.nbytes + h.node_off.nbytes + 4 * perm.size

    def _structure(self):
        store = self.model._graph_store()
        bs = store.assemble_device(self.h, self.packed, self.node_off, nnz_capacity=self.nnz_cap)
        bs.set_pooling(self.model.graph_pooling_type, self.dev)
        return bs

    def forward(self):
        model = self.model
        if self.fwd_graph is None:
            g = torch.cuda.CUDAGraph()
            params = [p.detach() for p in _engine.flat_params(model)]
            n0 = _ops.kernels_recorded()
            with quiet_capture(), torch.cuda.graph(g, pool=self.pool):
                bs = self._structure()
                g_f, d_logit, sv = _engine.run_forward(model, bs, self.neg_idx, True, True, None, params, self.comm)
            self.fwd_launches = _ops.kernels_recorded() - n0   # libgnm kernels inside the graph (C-side counters)
            _ops.REPLAYED[0] -= self.fwd_launches              # capture records, replay launches
            self.fwd_graph, self.g_f, self.d_logit, self.sv, self.params = g, g_f, d_logit, sv, params
            self.bwd_graph = None
        self.fwd_graph.replay()
        _ops.REPLAYED[0] += self.fwd_launches
        self.generation += 1
        return self.g_f.clone(), self.d_logit.clone()

    def backward(self, dg_f, dd_logit):
        if self.bwd_graph is None:
            self.dg_in = torch.zeros_like(self.g_f)
            self.dd_in = torch.zeros_like(self.d_logit)
            self.dg_in.copy_(dg_f)
            self.dd_in.copy_(dd_logit)
            g = torch.cuda.CUDAGraph()
            n0 = _ops.kernels_recorded()
            with quiet_capture(), torch.cuda.graph(g, pool=self.pool):
                _, grads = _engine.run_backward(self.model, self.sv, self.params, self.dg_in, self.dd_in, False,
                                                self.comm)
                self.grad_shapes = [None if x is None else tuple(x.shape) for x in grads]
                self.grad_sizes = [int(x.numel()) for x in grads if x is not None]
                self.flat = torch.cat([x.reshape(-1) for x in grads if x is not None])
            self.bwd_launches = _ops.kernels_recorded() - n0
            _ops.REPLAYED[0] -= self.bwd_launches
            self.bwd_graph = g
        else:
            self.dg_in.copy_(dg_f)
            self.dd_in.copy_(dd_logit)
        self.bwd_graph.replay()
        _ops.REPLAYED[0] += self.bwd_launches
        pieces = iter(self.flat.clone().split(self.grad_sizes))        # one call: the per-tensor slicing was host time
        return [None if shp is None else next(pieces).view(shp) for shp in self.grad_shapes]


class GraphedGINFunction(torch.autograd.Function):
    """Same contract as engine.GINFunction, executed by replaying the plan's CUDA graphs."""

    @staticmethod
    def forward(ctx, plan, *params):
        g_f, d_logit = plan.forward()
        ctx.plan = plan
        ctx.generation = plan.generation
        ctx.params = params
        return g_f, d_logit

    @staticmethod
    def backward(ctx, dg_f, dd_logit):
        plan = ctx.plan
        if ctx.generation != plan.generation:
            raise RuntimeError("GIN_InfoMaxReg: backward() of an earlier forward was called after a newer forward of the "
                               "same shape; with CUDA graphs the activations live in static buffers. Call backward "
                               "before the next forward, or set model.use_cuda_graphs = False.")
        if dg_f is None:
            dg_f = torch.zeros_like(plan.g_f)
        if dd_logit is None:
            dd_logit = torch.zeros_like(plan.d_logit)
        grads = plan.backward(dg_f, dd_logit)
        out = [g if (g is not None and p.requires_grad) else None for p, g in zip(ctx.params, grads)]
        return (None,) + tuple(out)
