"""Micro-benchmark / ncu target for the Linear kernels at the benchmark shape (M = 409,600 rows, 64 -> 64)."""
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
x = torch.randn(m, 64, device=dev)
w = torch.randn(64, 64, device=dev) * 0.2
b = torch.randn(64, device=dev)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
y = torch.empty(m, 64, device=dev)
st = torch.zeros(128, dtype=torch.float64, device=dev)
for impl in impls:
    ops.set_linear_impl(impl)
    for _ in range(2):
        ops.linear(x, w, False, b, sc, sh, y, st)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.linear(x, w, False, b, sc, sh, y, st)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters)]
    t = float(np.median(ts))
    print("%-8s median %8.1f us -> %7.1f GB/s (%.3f of 6546), %.1f TFLOP/s fp32-equivalent, abort=%s"
          % ("tcgen05" if impl == 2 else "ffma", t, 8.0 * m * 64 / t / 1e3, 8.0 * m * 64 / t / 1e3 / 6546.2,
             2.0 * m * 64 * 64 / t / 1e6, ops.aggregate_tc_status()))
ops.set_linear_impl(0)
