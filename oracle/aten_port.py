"""TEST / BASELINE INFRASTRUCTURE ONLY - a CPU port of the reference's hot path on the SAME
ATen operators the reference dispatches to.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package. The product path (`graph_neural_mapping_b200/`) never does.

Why it exists: `/root/reference` is not present on the GPU box and ships no installable
package (no setup.py), so the CPU baseline that `bench.py` times beside the GPU numbers
cannot import it there. This module restates `GIN_InfoMaxReg.forward`
(models/graphcnn.py:194-251) functionally over a state_dict, calling exactly the operators
the reference's modules call - `torch.spmm` on an UNCOALESCED sparse COO `Adj_block`
(graphcnn.py:104,154,178), `F.linear`, `F.batch_norm`, `F.bilinear` (nn.Bilinear,
discriminator.py:8,28-29), `F.dropout` - so its CPU cost profile is the reference's
(73 % `aten::_trilinear`), unlike the dense fp64 oracle in gin_oracle.py.
`cpu_baseline.kind` is therefore "port". Pinned against the reference's own outputs by
tests/test_oracle_vs_golden.py::test_aten_port_matches_reference.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _adjacency(graphs, learn_eps):
    """graphcnn.py:84-106."""
    parts, start = [], 0
    for g in graphs:
        parts.append(g.edge_mat + start)
        start += len(g.g)
    idx = torch.cat(parts, 1)
    val = torch.ones(idx.shape[1])
    if not learn_eps:
        loops = torch.arange(start).unsqueeze(0).repeat(2, 1)
        idx = torch.cat([idx, loops], 1)
        val = torch.cat([val, torch.ones(start)], 0)
    return torch.sparse_coo_tensor(idx, val, (start, start), check_invariants=False)


def _graph_pool(graphs, pooling):
    """graphcnn.py:109-134."""
    counts = [len(g.g) for g in graphs]
    rows = torch.repeat_interleave(torch.arange(len(graphs)), torch.tensor(counts))
    cols = torch.arange(sum(counts))
    if pooling == "average":
        val = torch.repeat_interleave(1.0 / torch.tensor(counts, dtype=torch.float32), torch.tensor(counts))
    else:
        val = torch.ones(sum(counts))
    return torch.sparse_coo_tensor(torch.stack([rows, cols]), val, (len(graphs), sum(counts)), check_invariants=False)


def _bn(x, sd, prefix, training):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training, 0.1, 1e-5)


def _mlp(x, sd, layer, k, training):
    """mlp.py:40-49."""
    p = "mlps.%d" % layer
    if k == 1:
        return F.linear(x, sd[p + ".linear.weight"], sd[p + ".linear.bias"])
    h = x
    for j in range(k - 1):
        h = F.relu(_bn(F.linear(h, sd["%s.linears.%d.weight" % (p, j)], sd["%s.linears.%d.bias" % (p, j)]), sd,
                       "%s.batch_norms.%d" % (p, j), training))
    return F.linear(h, sd["%s.linears.%d.weight" % (p, k - 1)], sd["%s.linears.%d.bias" % (p, k - 1)])


def forward(sd, graphs, cfg, training, final_dropout=0.0):
    """graphcnn.py:194-251 for sum/average neighbour pooling. `sd` maps state_dict keys to tensors
    (leaf tensors requiring grad for a training step; BN running buffers are updated in place).
    Draws `np.random.permutation(len(graphs))` exactly like graphcnn.py:199."""
    x = torch.cat([g.node_features for g in graphs], 0)
    pool = _graph_pool(graphs, cfg["graph_pooling_type"])
    idx = np.repeat(np.random.permutation(len(graphs)), len(graphs[0].node_features))     # :198-201
    adj = _adjacency(graphs, cfg["learn_eps"])
    hidden, h = [], x
    for layer in range(cfg["num_layers"]):
        pooled = torch.spmm(adj, h)                                                        # :154 / :178
        if cfg["neighbor_pooling_type"] == "average":
            pooled = pooled / torch.spmm(adj, torch.ones((adj.shape[0], 1)))               # :155-158
        if cfg["learn_eps"]:
            pooled = pooled + (1 + sd["eps"][layer]) * h                                   # :161
        rep = _mlp(pooled, sd, layer, cfg["num_mlp_layers"], training)
        h = F.relu(_bn(rep, sd, "batch_norms.%d" % layer, training))                       # :163-166
        hidden.append(h)
    c_logit, latent = 0, []
    for layer, h in enumerate(hidden):                                                     # :228-231
        pooled_h = torch.spmm(pool, h)
        c_logit = c_logit + F.dropout(F.linear(pooled_h, sd["linears_prediction.%d.weight" % layer],
                                               sd["linears_prediction.%d.bias" % layer]), final_dropout, training)
        latent.append(pooled_h)
    n_f, g_f = torch.cat(hidden, 1), torch.cat(latent, 1)
    c = torch.sigmoid(g_f)
    shuf = n_f[idx, :]                                                                     # :241-242
    reps = n_f.shape[0] // c.shape[0]
    c_x = torch.cat([row.expand(reps, n_f.shape[1]) for row in c], 0)                      # discriminator.py:23-26
    w, b = sd["disc.f_k.weight"], sd["disc.f_k.bias"]
    d_logit = torch.cat((F.bilinear(n_f, c_x, w, b), F.bilinear(shuf, c_x, w, b)), 0)      # discriminator.py:28-36
    return c_logit, d_logit, g_f


class TrainState(object):
    """Leaf parameters + buffers + Adam, stepping like main.py:25-41."""

    def __init__(self, state_dict, lr=0.005):
        self.sd = {}
        for k, v in state_dict.items():
            t = v.detach().clone().cpu()
            if t.is_floating_point() and "running_" not in k:
                t.requires_grad_(True)
            self.sd[k] = t
        self.params = [t for t in self.sd.values() if t.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=lr)

    def step(self, graphs, cfg, beta=0.05, final_dropout=0.0):
        c_logit, d_logit, _ = forward(self.sd, graphs, cfg, True, final_dropout)
        labels = torch.LongTensor([g.label for g in graphs])
        n = len(graphs) * graphs[0].node_features.shape[1]
        d_labels = torch.cat([torch.ones(n, 1), torch.zeros(n, 1)], 0)
        loss = F.cross_entropy(c_logit, labels) + beta * F.binary_cross_entropy_with_logits(d_logit, d_labels)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())
