"""TEST INFRASTRUCTURE ONLY - CPU oracle for the floating-point half of the hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package. The product path (`graph_neural_mapping_b200/`) never does.

A dense, dtype-parametrised (fp32 or fp64) restatement of the reference's GIN + DGI
arithmetic. The reference's arithmetic lives in PyTorch ATen (unpinned, `README.md:26`
`pytorch >= 1.4.0`; this image pins torch 2.11.0), so the restatement is written with
plain dense torch CPU tensor algebra - no `torch.spmm`, no `nn.Linear`, no
`nn.BatchNorm1d`, no `nn.Bilinear` - and differentiated by autograd on those primitives.
It exists because the reference cannot run in fp64 as written
(`torch.sparse.FloatTensor` + `torch.ones` force fp32, `graphcnn.py:93,104`): the fp64 run
is the tie-breaker that calibrates tolerances.

Parity pin: the reference ships no golden vectors. `tests/golden/make_golden.py`
imports the reference UNCHANGED from /root/reference, runs it on seeded synthetic graphs and
commits inputs + outputs as fixtures; `tests/test_oracle_vs_golden.py` checks this file
against every one of them.

Each function cites the reference lines it follows.
"""
import numpy as np
import torch

from . import csr_oracle

BN_EPS = 1e-5       # nn.BatchNorm1d default (mlp.py:38, graphcnn.py:51)
BN_MOMENTUM = 0.1


class OracleConfig(object):
    def __init__(self, num_layers, num_mlp_layers, input_dim, hidden_dim, output_dim,
                 final_dropout, learn_eps, graph_pooling_type, neighbor_pooling_type):
        self.num_layers = num_layers
        self.num_mlp_layers = num_mlp_layers
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.final_dropout = final_dropout
        self.learn_eps = learn_eps
        self.graph_pooling_type = graph_pooling_type
        self.neighbor_pooling_type = neighbor_pooling_type


def _batch_norm(x, prefix, params, buffers, training, new_buffers):
    """nn.BatchNorm1d forward (mlp.py:48, graphcnn.py:163,187): biased variance for the
    normalisation, unbiased for the running update, momentum 0.1."""
    w, b = params[prefix + ".weight"], params[prefix + ".bias"]
    if training:
        mean = x.mean(0)
        var = ((x - mean) ** 2).mean(0)
        n = x.shape[0]
        if new_buffers is not None:
            rm, rv = buffers[prefix + ".running_mean"], buffers[prefix + ".running_var"]
            unbiased = var.detach() * (n / max(n - 1, 1))
            new_buffers[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
            new_buffers[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * unbiased
            new_buffers[prefix + ".num_batches_tracked"] = buffers[prefix + ".num_batches_tracked"] + 1
    else:
        mean = buffers[prefix + ".running_mean"]
        var = buffers[prefix + ".running_var"]
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


def _mlp(x, layer, cfg, params, buffers, training, new_buffers):
    """mlp.py:40-49."""
    p = "mlps.%d" % layer
    if cfg.num_mlp_layers == 1:
        return x @ params[p + ".linear.weight"].t() + params[p + ".linear.bias"]
    h = x
    for k in range(cfg.num_mlp_layers - 1):
        z = h @ params["%s.linears.%d.weight" % (p, k)].t() + params["%s.linears.%d.bias" % (p, k)]
        h = torch.relu(_batch_norm(z, "%s.batch_norms.%d" % (p, k), params, buffers, training, new_buffers))
    k = cfg.num_mlp_layers - 1
    return h @ params["%s.linears.%d.weight" % (p, k)].t() + params["%s.linears.%d.bias" % (p, k)]


def _maxpool_neighbors(h, padded, learn_eps):
    """graphcnn.py:137-143: append the column-wise minimum as a dummy row (index -1 of the
    padded list hits it), gather the padded neighbour list and take the max over it. The
    gather + max are kept in the reference's candidate order so that ties (the one-hot
    input has many) send the gradient to the same entry."""
    dummy = h.min(dim=0)[0]
    h_with_dummy = torch.cat([h, dummy.reshape(1, -1)], 0)
    return h_with_dummy[padded].max(dim=1)[0]


def padded_neighbor_list(graphs, learn_eps):
    """graphcnn.py:55-81."""
    max_deg = max(g.max_neighbor for g in graphs)
    rows = []
    start = 0
    for g in graphs:
        for j, nb in enumerate(g.neighbors):
            pad = [n + start for n in nb]
            pad.extend([-1] * (max_deg - len(pad)))
            if not learn_eps:
                pad.append(j + start)
            rows.append(pad)
        start += len(g.g)
    return torch.tensor(rows, dtype=torch.long)


def _layer(h, layer, adj, deg, cfg, params, buffers, training, new_buffers, neighbor_lists=None):
    """graphcnn.py:146-167 (`next_layer_eps`) and :170-191 (`next_layer`)."""
    if cfg.neighbor_pooling_type == "max":
        pooled = _maxpool_neighbors(h, neighbor_lists, cfg.learn_eps)
    else:
        pooled = adj @ h                                       # :154 / :178
        if cfg.neighbor_pooling_type == "average":
            pooled = pooled / deg                              # :155-158 / :179-182 (0/0 -> NaN kept)
    if cfg.learn_eps:
        pooled = pooled + (1 + params["eps"][layer]) * h       # :161
    rep = _mlp(pooled, layer, cfg, params, buffers, training, new_buffers)   # :162 / :185
    h = _batch_norm(rep, "batch_norms.%d" % layer, params, buffers, training, new_buffers)  # :163 / :187
    return torch.relu(h)                                       # :166 / :190


def _prepare(graphs, cfg, dtype):
    node_counts = [len(g.g) for g in graphs]
    edge_mats = [g.edge_mat.numpy() for g in graphs]
    adj = torch.from_numpy(csr_oracle.dense_adjacency(edge_mats, node_counts, cfg.learn_eps)).to(dtype)
    deg = adj.sum(1, keepdim=True)                             # spmm(Adj_block, ones) :157,181
    off, scale = csr_oracle.graph_pool_segments(node_counts, cfg.graph_pooling_type)
    m = int(off[-1])
    pool = torch.zeros(len(graphs), m, dtype=dtype)
    for i in range(len(graphs)):
        pool[i, off[i]:off[i + 1]] = scale[i]                  # graphcnn.py:120-129
    x = torch.cat([g.node_features for g in graphs], 0).to(dtype)   # :195
    neighbor_lists = None
    if cfg.neighbor_pooling_type == "max":
        neighbor_lists = padded_neighbor_list(graphs, cfg.learn_eps)
    return x, adj, deg, pool, neighbor_lists


def forward(state, graphs, perm, cfg, training, dtype=torch.float64, x_override=None,
            with_dgi=True):
    """graphcnn.py:194-251. `state` maps state_dict keys to tensors (params may require
    grad). `perm` is the value `np.random.permutation(len(batch_graph))` returned at
    `graphcnn.py:199`. Dropout (`:230`) is only restated for p == 0 or eval mode.

    Returns dict(c_logit, d_logit, g_f, n_f, new_buffers).
    """
    assert (not training) or cfg.final_dropout == 0.0, "oracle restates dropout only as identity"
    params = {k: v for k, v in state.items()}
    buffers = params
    new_buffers = {} if training else None
    x, adj, deg, pool, neighbor_lists = _prepare(graphs, cfg, dtype)
    if x_override is not None:
        x = x_override
    hidden = []
    h = x
    for layer in range(cfg.num_layers):                        # :212-222
        h = _layer(h, layer, adj, deg, cfg, params, buffers, training, new_buffers, neighbor_lists)
        hidden.append(h)
    c_logit = 0
    latent = []
    for layer, h in enumerate(hidden):                         # :228-231
        pooled_h = pool @ h
        w = params["linears_prediction.%d.weight" % layer]
        b = params["linears_prediction.%d.bias" % layer]
        c_logit = c_logit + (pooled_h @ w.t() + b)
        latent.append(pooled_h)
    n_f = torch.cat(hidden, 1)                                 # :233
    g_f = torch.cat(latent, 1)                                 # :234
    out = dict(c_logit=c_logit, g_f=g_f, n_f=n_f, new_buffers=new_buffers, hidden=hidden, x=x)
    if with_dgi:
        c = torch.sigmoid(g_f)                                 # :238-239
        n_first = graphs[0].node_features.shape[0]
        idx = torch.from_numpy(csr_oracle.dgi_negative_rows(perm, n_first))   # :198-201
        shuf = n_f[idx, :]                                     # :241-242
        out["d_logit"] = discriminator(params["disc.f_k.weight"], params["disc.f_k.bias"], c, n_f, shuf)  # :246
    return out


def discriminator(weight, bias, c, h_pl, h_mi, s_bias1=None, s_bias2=None):
    """discriminator.py:19-38 with nn.Bilinear(n_h, n_h, 1) written out:
    f_k(x1, x2) = x1^T W[0] x2 + b."""
    reps = h_pl.shape[0] // c.shape[0]                         # :23-26
    c_x = c.repeat_interleave(reps, dim=0)
    w = weight[0]
    sc_1 = ((h_pl @ w) * c_x).sum(1, keepdim=True) + bias      # :28
    sc_2 = ((h_mi @ w) * c_x).sum(1, keepdim=True) + bias      # :29
    if s_bias1 is not None:
        sc_1 = sc_1 + s_bias1
    if s_bias2 is not None:
        sc_2 = sc_2 + s_bias2
    return torch.cat((sc_1, sc_2), 0)                          # :36


def losses(c_logit, d_logit, labels, n_graphs, num_rois, beta):
    """main.py:16-17,31-37: CrossEntropy(c_logit, labels) + beta * BCEWithLogits(d_logit,
    [ones(B*num_rois); zeros(B*num_rois)])."""
    labels = torch.as_tensor(labels, dtype=torch.long)
    logp = c_logit - torch.logsumexp(c_logit, dim=1, keepdim=True)
    c_loss = -logp[torch.arange(c_logit.shape[0]), labels].mean()
    d_labels = torch.cat([torch.ones(n_graphs * num_rois, 1), torch.zeros(n_graphs * num_rois, 1)], 0).to(d_logit.dtype)
    x = d_logit
    d_loss = (torch.clamp(x, min=0) - x * d_labels + torch.log1p(torch.exp(-x.abs()))).mean()
    return c_loss + beta * d_loss, c_loss, d_loss


def train_step_grads(state_dict, graphs, perm, cfg, beta, dtype=torch.float64):
    """One `main.py:25-41` step without the optimiser: forward (train mode), loss,
    backward. Returns dict(c_logit, d_logit, g_f, loss, grads{name: tensor|None}, new_buffers)."""
    state = {}
    for k, v in state_dict.items():
        t = v.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype)
            if "running_" not in k:
                t.requires_grad_(True)
        state[k] = t
    out = forward(state, graphs, perm, cfg, training=True, dtype=dtype)
    labels = [g.label for g in graphs]
    num_rois = graphs[0].node_features.shape[1]                # main.py:22
    loss, c_loss, d_loss = losses(out["c_logit"], out["d_logit"], labels, len(graphs), num_rois, beta)
    names = [k for k, v in state.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [state[k] for k in names], allow_unused=True)
    return dict(c_logit=out["c_logit"].detach(), d_logit=out["d_logit"].detach(), g_f=out["g_f"].detach(),
                n_f=out["n_f"].detach(), loss=loss.detach(), c_loss=c_loss.detach(), d_loss=d_loss.detach(),
                grads=dict(zip(names, grads)), new_buffers=out["new_buffers"])


def saliency(state_dict, graphs, cls, cfg, dtype=torch.float64):
    """graphcnn.py:254-299: eval-mode forward without DGI, backward of
    score . onehot(cls) to X_concat. Batched input is allowed here (exact)."""
    state = {}
    for k, v in state_dict.items():
        t = v.detach().clone()
        state[k] = t.to(dtype) if t.is_floating_point() else t
    x = torch.cat([g.node_features for g in graphs], 0).to(dtype).requires_grad_(True)
    out = forward(state, graphs, None, cfg, training=False, dtype=dtype, x_override=x, with_dgi=False)
    score = out["c_logit"]
    cot = torch.zeros_like(score)                              # :263-264 (per graph row)
    cot[:, cls] = 1
    (g,) = torch.autograd.grad(score, x, cot)
    return g.detach(), score.detach()
