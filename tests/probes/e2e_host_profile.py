"""Host-side profile (cProfile) of the literal main.py loop body against the B200 model: where the CPU time of one
e2e step goes while the GPU is idle. usage: e2e_host_profile.py [steps]"""
import cProfile, os, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graph_neural_mapping_b200 import synth
from graph_neural_mapping_b200.models import GIN_InfoMaxReg

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
B, N = 1024, 400
pool = synth.make_graphs_bulk(B, N, 30, 256, seed0=0, device=dev)
model = GIN_InfoMaxReg(5, 2, N, 64, 2, 0.5, False, "sum", "sum", dev).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.01)
c_crit, d_crit = torch.nn.CrossEntropyLoss(), torch.nn.BCEWithLogitsLoss()
model.train()
seg = {"model": 0.0, "labels": 0.0, "loss+bwd": 0.0, "opt": 0.0, "readback": 0.0}


def step():
    t0 = time.perf_counter()
    sel = np.random.permutation(len(pool))[:B]
    batch = [pool[i] for i in sel]
    c_logit, d_logit = model(batch)
    t1 = time.perf_counter()
    c_labels = torch.LongTensor([g.label for g in batch]).to(dev)
    d_labels = torch.cat([torch.ones(B * N, 1), torch.zeros(B * N, 1)], 0).to(dev)
    t2 = time.perf_counter()
    loss = c_crit(c_logit, c_labels) + 0.1 * d_crit(d_logit, d_labels)
    opt.zero_grad()
    loss.backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    v = float(loss.detach().cpu().numpy())
    t5 = time.perf_counter()
    for k, d in zip(seg, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        seg[k] += d
    return v


for _ in range(5):
    step()
for k in seg:
    seg[k] = 0.0
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
pr.disable()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("e2e %.2f ms/step -> %.0f graphs/s" % (1e3 * dt / steps, B * steps / dt))
print({k: "%.2f ms" % (1e3 * v / steps) for k, v in seg.items()})
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
