"""Per-kernel time table of one batched saliency call (probe)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from graph_neural_mapping_b200 import ops, synth
from graph_neural_mapping_b200.models import GIN_InfoMaxReg

dev = torch.device("cuda")
b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pool = synth.make_graphs_bulk(b, 400, 30, 128, seed0=0, device=dev)
model = GIN_InfoMaxReg(5, 2, 400, 64, 2, 0.5, False, "sum", "sum", dev).to(dev)
for _ in range(3):
    model.compute_saliency_batched(pool, 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
model.compute_saliency_batched(pool, 1)
torch.cuda.synchronize()
print("wall %.2f ms" % ((time.perf_counter() - t0) * 1e3))
with bench.OpTimer(ops) as t:
    model.compute_saliency_batched(pool, 1)
    tab = t.table()
for k, v in sorted(tab.items(), key=lambda kv: -kv[1][1]):
    print("%-36s %4d %9.3f ms" % (k, v[0], v[1]))
print("sum %.3f ms" % sum(v[1] for v in tab.values()))
for i in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sal = model.compute_saliency_batched(pool, 1)
    torch.cuda.synchronize()
    print("call %d wall %.2f ms  alloc %.0f MB reserved %.0f MB" % (i, (time.perf_counter() - t0) * 1e3,
          torch.cuda.memory_allocated() / 1e6, torch.cuda.memory_reserved() / 1e6))
