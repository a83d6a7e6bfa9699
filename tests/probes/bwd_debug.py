import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graph_neural_mapping_b200 import ops
dev = torch.device("cuda")
m = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
act = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
dy, z, x = torch.randn(m, 64, device=dev), torch.randn(m, 64, device=dev), torch.randn(m, 64, device=dev)
w = torch.randn(64, 64, device=dev) * 0.2
coef = torch.randn(3, 64, device=dev)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
mu, rs = torch.randn(64, device=dev), torch.rand(64, device=dev) + 0.5
res = {}
for impl in (2, 1):
    ops.set_linear_impl(impl)
    dx = torch.full((m, 64), float("nan"), device=dev)
    dw, db = torch.zeros(64, 64, device=dev), torch.zeros(64, device=dev)
    st = torch.zeros(128, dtype=torch.float64, device=dev)
    if act:
        ops.linear_bwd(dy, z, coef, x, sc, sh, mu, rs, w, dw, db, dx, st)
    else:
        ops.linear_bwd(dy, z, coef, x, None, None, None, None, w, dw, db, dx, None)
    res[impl] = (dx, dw, db, st)
ops.set_linear_impl(0)
print("abort", ops.aggregate_tc_status())
d = (res[2][0] - res[1][0]).abs()
rowerr = d.max(1).values
bad = (rowerr > 1e-3).nonzero().flatten().cpu().numpy()
print("bad rows", bad.size, "of", m)
if bad.size:
    tiles = np.unique(bad // 128)
    print("bad tiles", tiles[:40], "count", tiles.size)
    print("tile -> it", [(int(t), int(t) // 148) for t in tiles[:20]])
    r = int(bad[0])
    print("row", r, "tc", res[2][0][r, :8].tolist(), "ff", res[1][0][r, :8].tolist())
    cols = (d[bad] > 1e-3).any(0).nonzero().flatten().tolist()
    print("bad cols", cols)
print("dw err", float((res[2][1] - res[1][1]).abs().max()), "scale", float(res[1][1].abs().max()))
print("db err", float((res[2][2] - res[1][2]).abs().max()), "scale", float(res[1][2].abs().max()))
print("st err", float((res[2][3] - res[1][3]).abs().max()), "scale", float(res[1][3].abs().max()))
