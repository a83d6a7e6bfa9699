"""Micro-benchmark / ncu target for the backward unit at the benchmark shape (M = 409,600 rows, 64 -> 64).
usage: bwd_bench.py [rows] [iters] [impl ...]   impl 2 = tcgen05 (one-pass kernel when dx is requested), 3 = tcgen05 two-pass
pair (dx kernel + wgrad kernel), 1 = fused FFMA kernel"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graph_neural_mapping_b200 import ops

m = int(sys.argv[1]) if len(sys.argv) > 1 else 409600
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
impls = [int(v) for v in sys.argv[3:]] or [2, 1]
dev = torch.device("cuda")
torch.manual_seed(0)
dy, z, x = torch.randn(m, 64, device=dev), torch.randn(m, 64, device=dev), torch.randn(m, 64, device=dev)
w = torch.randn(64, 64, device=dev) * 0.2
coef = torch.randn(3, 64, device=dev)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
mu, rs = torch.randn(64, device=dev), torch.rand(64, device=dev) + 0.5
dx = torch.empty(m, 64, device=dev)
dw, db = torch.zeros(64, 64, device=dev), torch.zeros(64, device=dev)
st = torch.zeros(128, dtype=torch.float64, device=dev)
for impl in impls:
    ops.set_linear_impl(impl)
    for with_dx in (True, False):
        args = (dy, z, coef, x, sc, sh, mu, rs, w, dw, db, dx if with_dx else None, st if with_dx else None)
        for _ in range(2):
            ops.linear_bwd(*args)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        ev[0].record()
        for i in range(iters):
            ops.linear_bwd(*args)
            ev[i + 1].record()
        torch.cuda.synchronize()
        t = float(np.median([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters)]))
        nbytes = (16.0 if with_dx else 12.0) * m * 64          # dy, z, x read once (+ dx written)
        print("%-8s %-8s median %8.1f us -> %7.1f GB/s algorithmic (%.3f of 6546), abort=%s"
              % ({2: "tcgen05", 3: "tc-2pass"}.get(impl, "ffma"), "dx+dw" if with_dx else "dw only", t, nbytes / t / 1e3,
                 nbytes / t / 1e3 / 6546.2, ops.aggregate_tc_status()))
ops.set_linear_impl(0)
