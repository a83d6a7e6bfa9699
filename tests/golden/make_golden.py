#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

For each case it builds seeded synthetic graphs (`graph_neural_mapping_b200/synth.py`, routed
through the literal networkx construction of `util.py:43-103`), constructs the reference's
`GIN_InfoMaxReg` (`/root/reference/models/graphcnn.py:12`), replays one `main.py:25-41`
training step (forward in train mode, CE + beta*BCE loss, backward), one eval forward,
the latent extraction and `compute_saliency`, and stores inputs + outputs in
`<case>.npz`. The reference ships no golden vectors of its own, so these files
are the parity pin for both the oracle (`oracle/`) and the CUDA path.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.path.insert(0, REPO)
from graph_neural_mapping_b200 import synth  # noqa: E402

# import the reference exactly as its main.py does (cwd matters for its "models/" append)
os.chdir(REF)
sys.path.insert(0, REF)
for _m in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
    del sys.modules[_m]
warnings.filterwarnings("ignore")
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("ref_graphcnn", os.path.join(REF, "models", "graphcnn.py"))
ref_graphcnn = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_graphcnn)
RefModel = ref_graphcnn.GIN_InfoMaxReg

c_criterion = torch.nn.CrossEntropyLoss()      # main.py:16
d_criterion = torch.nn.BCEWithLogitsLoss()     # main.py:17

CASES = [
    # name, graphs(B, N, n_time, seed0), model cfg
    dict(name="tiny_eps_sum", B=4, N=12, T=64, seed0=100,
         cfg=dict(num_layers=3, num_mlp_layers=2, hidden_dim=8, learn_eps=True, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    dict(name="tiny_noeps_sum", B=4, N=12, T=64, seed0=100,
         cfg=dict(num_layers=3, num_mlp_layers=2, hidden_dim=8, learn_eps=False, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    dict(name="tiny_eps_avg", B=5, N=16, T=64, seed0=300, need_no_isolated=True,
         cfg=dict(num_layers=2, num_mlp_layers=2, hidden_dim=12, learn_eps=True, graph_pooling_type="average", neighbor_pooling_type="average")),
    dict(name="tiny_noeps_avg", B=3, N=16, T=64, seed0=300,
         cfg=dict(num_layers=2, num_mlp_layers=2, hidden_dim=12, learn_eps=False, graph_pooling_type="average", neighbor_pooling_type="average")),
    dict(name="tiny_mlp1", B=3, N=10, T=64, seed0=500,
         cfg=dict(num_layers=2, num_mlp_layers=1, hidden_dim=8, learn_eps=True, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    dict(name="tiny_mlp3", B=3, N=10, T=64, seed0=500,
         cfg=dict(num_layers=2, num_mlp_layers=3, hidden_dim=8, learn_eps=False, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    dict(name="tiny_eps_max", B=3, N=12, T=64, seed0=700,
         cfg=dict(num_layers=2, num_mlp_layers=2, hidden_dim=8, learn_eps=True, graph_pooling_type="sum", neighbor_pooling_type="max")),
    dict(name="tiny_noeps_max", B=3, N=12, T=64, seed0=700,
         cfg=dict(num_layers=2, num_mlp_layers=2, hidden_dim=8, learn_eps=False, graph_pooling_type="average", neighbor_pooling_type="max")),
    dict(name="mid_eps_sum_h64", B=6, N=48, T=128, seed0=900,
         cfg=dict(num_layers=5, num_mlp_layers=2, hidden_dim=64, learn_eps=True, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    dict(name="schaefer400_noeps", B=3, N=400, T=1200, seed0=0, light=True,
         cfg=dict(num_layers=5, num_mlp_layers=2, hidden_dim=64, learn_eps=False, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    dict(name="schaefer400_eps", B=2, N=400, T=1200, seed0=10, light=True,
         cfg=dict(num_layers=5, num_mlp_layers=2, hidden_dim=64, learn_eps=True, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    # M = B*N = 6400 rows >= 4096: the row count from which libgnm takes its tcgen05 GEMM family - the code path the
    # benchmark runs (BASELINE configs[1] shape class). B = 16 also shards over 2, 4 and 8 ranks for the data-parallel
    # parity runs. `compact`: the graphs keep synth.make_graph's canonical edge order (upper-triangle pairs row-major,
    # then the same pairs reversed, util.py:99-103) and are stored as adjacency bitmaps instead of int64 edge lists.
    dict(name="schaefer400_b16_noeps", B=16, N=400, T=1200, seed0=2000, light=True, compact=True,
         cfg=dict(num_layers=5, num_mlp_layers=2, hidden_dim=64, learn_eps=False, graph_pooling_type="sum", neighbor_pooling_type="sum")),
    dict(name="schaefer400_b16_eps", B=16, N=400, T=1200, seed0=2100, light=True, compact=True,
         cfg=dict(num_layers=5, num_mlp_layers=2, hidden_dim=64, learn_eps=True, graph_pooling_type="sum", neighbor_pooling_type="sum")),
]
BETA = 0.05        # main.py:118 default


def has_isolated(g):
    n = len(g.g)
    deg = np.bincount(g.edge_mat.numpy()[0], minlength=n)
    return bool((deg == 0).any())


def build_graphs(case):
    graphs = []
    seed = case["seed0"]
    while len(graphs) < case["B"]:
        g = synth.make_graph(seed, case["N"], 30, case["T"])
        seed += 1
        if case.get("need_no_isolated") and has_isolated(g):
            continue
        graphs.append(g if case.get("compact") else synth.to_networkx_route(g))
    return graphs


def run_case(case):
    cfg = dict(case["cfg"])
    graphs = build_graphs(case)
    N = case["N"]
    torch.manual_seed(1000 + case["seed0"])
    model = RefModel(cfg["num_layers"], cfg["num_mlp_layers"], N, cfg["hidden_dim"], 2, 0.0,
                     cfg["learn_eps"], cfg["graph_pooling_type"], cfg["neighbor_pooling_type"], torch.device("cpu"))
    # make the test non-trivial: non-zero eps, non-default BN affine/running stats, non-zero disc bias
    with torch.no_grad():
        model.eps.copy_(torch.linspace(-0.3, 0.4, cfg["num_layers"]))
        model.disc.f_k.bias.fill_(0.05)
        if case.get("light"):
            # sum pooling over 400 nodes saturates the 2-class softmax at random init (CE gradient
            # exactly 0 in fp32); shrink the heads so the fixture exercises every gradient path
            for lin in model.linears_prediction:
                lin.weight.mul_(0.01)
        for name, p in model.named_parameters():
            if "batch_norms" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn_like(p))
            if "batch_norms" in name and name.endswith("bias"):
                p.copy_(0.1 * torch.randn_like(p))
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(0.3 * torch.randn_like(b))
            if name.endswith("running_var"):
                b.copy_(0.5 + torch.rand_like(b))
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}

    out = {}
    out["config"] = np.array(json.dumps(dict(cfg, input_dim=N, output_dim=2, final_dropout=0.0, beta=BETA, N=N, B=case["B"])))
    em = [g.edge_mat.contiguous().numpy() for g in graphs]
    if case.get("compact"):
        bits = np.zeros((len(graphs), N, N), dtype=bool)
        for i, e in enumerate(em):
            half = e.shape[1] // 2
            assert np.array_equal(e[:, half:], e[::-1, :half]) and (e[0, :half] < e[1, :half]).all()
            bits[i, e[0, :half], e[1, :half]] = True
            iu, ju = np.nonzero(bits[i])
            assert np.array_equal(iu, e[0, :half]) and np.array_equal(ju, e[1, :half])      # canonical order round-trips
        out["edge_bits"] = np.packbits(bits.reshape(len(graphs), -1), axis=1)
    else:
        out["edge_cat"] = np.concatenate(em, axis=1)
        out["edge_off"] = np.cumsum([0] + [e.shape[1] for e in em]).astype(np.int64)
    out["node_counts"] = np.array([len(g.g) for g in graphs], dtype=np.int64)
    out["labels"] = np.array([g.label for g in graphs], dtype=np.int64)
    for k, v in state0.items():
        out["state/" + k] = v.numpy()

    # ---- the sparse index objects the reference builds (graphcnn.py:84-134) ----
    if cfg["neighbor_pooling_type"] != "max" and not case.get("compact"):
        adj = model._GIN_InfoMaxReg__preprocess_neighbors_sumavepool(graphs)
        out["adj_coo_idx"] = adj._indices().numpy().copy()
        adjc = adj.coalesce()
        csr = adjc.to_sparse_csr()
        out["adj_crow"] = csr.crow_indices().numpy().copy()
        out["adj_col"] = csr.col_indices().numpy().copy()
        out["adj_val"] = csr.values().numpy().copy()
    pool = model._GIN_InfoMaxReg__preprocess_graphpool(graphs).coalesce()
    out["pool_idx"] = pool.indices().numpy().copy()
    out["pool_val"] = pool.values().numpy().copy()

    # ---- one training step: main.py:25-41 (no optimiser step) ----
    model.train()
    np.random.seed(4242 + case["seed0"])
    perm = np.random.permutation(len(graphs))
    np.random.seed(4242 + case["seed0"])
    c_logit, d_logit = model(graphs)
    out["perm"] = perm.astype(np.int64)
    c_labels = torch.LongTensor([g.label for g in graphs])
    num_rois = graphs[0].node_features.shape[1]
    d_labels = torch.cat([torch.ones(len(graphs) * num_rois, 1), torch.zeros(len(graphs) * num_rois, 1)], 0)
    d_loss = d_criterion(d_logit, d_labels)
    c_loss = c_criterion(c_logit, c_labels)
    loss = c_loss + BETA * d_loss
    model.zero_grad()
    loss.backward()
    out["train/c_logit"] = c_logit.detach().numpy()
    out["train/d_logit"] = d_logit.detach().numpy()
    out["train/loss"] = loss.detach().numpy()
    out["train/c_loss"] = c_loss.detach().numpy()
    out["train/d_loss"] = d_loss.detach().numpy()
    for name, p in model.named_parameters():
        if p.grad is None:
            out["gradnone/" + name] = np.array(1)
        else:
            out["grad/" + name] = p.grad.detach().numpy().copy()
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["buf_after/" + k] = v.numpy().copy()

    # ---- eval forward on the batch + latent (graphcnn.py:248-249), with the updated buffers ----
    model.eval()
    np.random.seed(777)
    perm_e = np.random.permutation(len(graphs))
    np.random.seed(777)
    with torch.no_grad():
        c_e, d_e = model(graphs)
    out["eval/perm"] = perm_e.astype(np.int64)
    out["eval/c_logit"] = c_e.numpy()
    out["eval/d_logit"] = d_e.numpy()
    out["eval/latent"] = model(graphs, latent=True)
    # per-graph eval (main.py:53-54) for the first graph
    np.random.seed(778)
    with torch.no_grad():
        c1, d1 = model([graphs[0]])
    out["eval1/c_logit"] = c1.numpy()
    out["eval1/d_logit"] = d1.numpy()

    # ---- saliency (graphcnn.py:254-299), graph 0 and 1, both classes ----
    n_sal = 1 if case.get("light") else 2
    for gi in range(n_sal):
        for cls in ([1] if case.get("light") else [0, 1]):
            sal = model.compute_saliency([graphs[gi]], cls)
            out["saliency/g%d_c%d" % (gi, cls)] = sal.detach().numpy().copy()
            if gi == 0 and cls == 1:
                for name, p in model.named_parameters():
                    if p.grad is not None and not case.get("light"):
                        out["saliency_paramgrad/" + name] = p.grad.detach().numpy().copy()
    path = os.path.join(HERE, case["name"] + ".npz")
    np.savez_compressed(path, **out)
    print("%-22s loss=%.6f  %7.1f KB" % (case["name"], float(loss), os.path.getsize(path) / 1024.0))


def main():
    only = sys.argv[1:]
    for case in CASES:
        if only and case["name"] not in only:
            continue
        run_case(case)


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
