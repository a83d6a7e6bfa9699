"""Synthetic HCP-shaped ROI connectivity graphs.

The reference ships no data (`.gitignore:106` ignores `data/`), so every test and
benchmark input is generated here, following the reference's own construction:

* `dataset.py:93-101` (`DataEdges.get_adjacency`): `mask = fc > percentile(fc, 100-sparsity)`
  over ALL N*N entries (diagonal included), then only `neighbor > node` pairs are kept.
* `util.py:43-103` (`load_data`): an undirected graph over nodes 0..N-1, `edge_mat` lists
  every edge in both directions, the second half being the first half reversed
  (`util.py:99-103`); `node_features` is one-hot over the ROI tag (`util.py:114-116`), i.e. the
  identity matrix for one subject's 400 unique ROIs; `label` is the class index.

`SynthGraph` carries exactly the fields `GIN_InfoMaxReg` reads from `S2VGraph`
(`util.py:9-17`): `g` (only `len(g)` is used), `label`, `node_features`, `edge_mat`,
`neighbors`, `max_neighbor`.
"""
from __future__ import annotations

import numpy as np
import torch


class _NodeSet(object):
    """Stand-in for the networkx graph: the model only ever calls `len(graph.g)`."""

    __slots__ = ("n",)

    def __init__(self, n):
        self.n = int(n)

    def __len__(self):
        return self.n


class SynthGraph(object):
    """Field-compatible with the reference's `S2VGraph` (`util.py:9-17`)."""

    def __init__(self, n_nodes, label, edge_mat, node_features, with_neighbors=False):
        self.label = int(label)
        self.g = _NodeSet(n_nodes)
        self.node_tags = list(range(n_nodes))
        self.node_features = node_features
        self.edge_mat = edge_mat
        self.neighbors = []
        self.max_neighbor = 0
        if with_neighbors:
            nb = [[] for _ in range(n_nodes)]
            half = edge_mat.shape[1] // 2
            em = edge_mat.numpy()
            for i, j in zip(em[0, :half].tolist(), em[1, :half].tolist()):
                nb[i].append(j)
                nb[j].append(i)
            self.neighbors = nb
            self.max_neighbor = max((len(x) for x in nb), default=0)


def connectivity_matrix(seed, n_rois=400, n_time=1200, n_factors=7):
    """A correlation matrix with community structure and a hub/low-degree spread.

    K shared network factors: iid noise alone gives a too-flat degree
    range. Counter-based: the seed alone determines the graph.
    """
    rng = np.random.default_rng([int(seed), int(n_rois), 0x5EED])
    member = rng.integers(0, n_factors, size=n_rois)
    load = np.zeros((n_rois, n_factors))
    load[np.arange(n_rois), member] = rng.uniform(0.4, 1.4, size=n_rois)
    load += rng.uniform(0.0, 0.25, size=(n_rois, n_factors))
    factors = rng.standard_normal((n_factors, n_time))
    ts = load @ factors + rng.standard_normal((n_rois, n_time))
    return np.corrcoef(ts)


def edges_from_connectivity(fc, sparsity=30):
    """`dataset.py:93-101` thresholding + `util.py:99-103` mirrored edge list.

    Returns int64 [2, E] with the upper-triangle pairs first (row-major) and the same
    pairs reversed second.
    """
    thr = np.percentile(fc, 100 - sparsity)
    mask = fc > thr
    iu, ju = np.nonzero(np.triu(mask, 1))
    src = np.concatenate([iu, ju]).astype(np.int64)
    dst = np.concatenate([ju, iu]).astype(np.int64)
    return np.stack([src, dst], 0)


def make_graph(seed, n_rois=400, sparsity=30, n_time=1200, label=None, with_neighbors=False):
    fc = connectivity_matrix(seed, n_rois, n_time)
    em = torch.from_numpy(edges_from_connectivity(fc, sparsity))
    feats = torch.eye(n_rois, dtype=torch.float32)
    return SynthGraph(n_rois, seed % 2 if label is None else label, em, feats, with_neighbors)


def make_graphs(n_graphs, n_rois=400, sparsity=30, n_time=1200, seed0=0, with_neighbors=False):
    return [make_graph(seed0 + i, n_rois, sparsity, n_time, with_neighbors=with_neighbors)
            for i in range(n_graphs)]


def make_graphs_bulk(n_graphs, n_rois=400, sparsity=30, n_time=256, seed0=0, device="cpu",
                     share_features=True, edges_on_device=False):
    """Bulk generator for throughput runs (same recipe, batched in torch on `device`).

    Not bit-identical to `make_graph` (different RNG stream); use `make_graph` where a
    graph must be reproducible across machines. `share_features=True` makes every graph
    reference one identity matrix (as one subject list would after `util.py:114-116`
    only in value, here also in storage) to keep host memory bounded.

    The whole recipe - time series, corrcoef, percentile threshold (`dataset.py:93-101`), upper-triangle edge
    extraction and the mirrored edge list (`util.py:99-103`) - runs on `device`, vectorised over chunks of 64 graphs
    (one `nonzero` per chunk, no per-graph work on the host). `edges_on_device=True` leaves each graph's `edge_mat` on
    the device as well: `GraphStore.ensure` ingests such graphs without a host round trip (needed for
    the 100k-graph saliency configuration, whose int64 edge lists alone are 76 GB).
    """
    gen = torch.Generator(device=device)
    gen.manual_seed(1234567 + seed0)
    n_factors = 7
    graphs = []
    eye = torch.eye(n_rois, dtype=torch.float32)
    chunk = 64
    k = int(np.floor((100 - sparsity) / 100.0 * (n_rois * n_rois - 1)))
    for c0 in range(0, n_graphs, chunk):
        c = min(chunk, n_graphs - c0)
        member = torch.randint(0, n_factors, (c, n_rois), generator=gen, device=device)
        load = torch.rand(c, n_rois, n_factors, generator=gen, device=device) * 0.25
        amp = torch.rand(c, n_rois, generator=gen, device=device) + 0.4
        load.scatter_add_(2, member.unsqueeze(-1), amp.unsqueeze(-1))
        factors = torch.randn(c, n_factors, n_time, generator=gen, device=device)
        ts = torch.bmm(load, factors) + torch.randn(c, n_rois, n_time, generator=gen, device=device)
        ts = ts - ts.mean(-1, keepdim=True)
        ts = ts / ts.norm(dim=-1, keepdim=True)
        fc = torch.bmm(ts, ts.transpose(1, 2))
        flat = fc.reshape(c, -1)
        # np.percentile's linear interpolation lies between order statistics k and k+1
        # (0-based, ascending); "fc > thr" therefore keeps exactly the entries ranked above k.
        thr = torch.kthvalue(flat, k + 1, dim=1).values
        mask = torch.triu(fc > thr.view(c, 1, 1), 1)
        # edge extraction on the device: one nonzero over the chunk (rows come out sorted by graph, then row-major
        # within the graph = the upper-triangle order of edges_from_connectivity), split by per-graph counts
        nz = torch.nonzero(mask)                                   # [E_chunk, 3]: graph, i, j
        counts = torch.bincount(nz[:, 0], minlength=c)
        offs = [0] + torch.cumsum(counts, 0).tolist()              # the only host synchronisation of the chunk
        ij = nz[:, 1:].t().contiguous()                            # [2, E_chunk]
        if not edges_on_device:
            ij = ij.cpu()
        for b in range(c):
            half = ij[:, offs[b]:offs[b + 1]]
            em = torch.cat([half, half.flip(0)], 1).contiguous()    # util.py:99-103: the same pairs, reversed, second
            feats = eye if share_features else eye.clone()
            graphs.append(SynthGraph(n_rois, (seed0 + c0 + b) % 2, em, feats))
    return graphs


def release_edges(graphs):
    """Drop the edge lists of graphs a GraphStore has already ingested (their device CSR / bitmap is what the kernels
    read). The graph objects stay valid cache keys; a store that evicts them cannot rebuild them."""
    empty = torch.zeros(2, 0, dtype=torch.int64)
    for g in graphs:
        g.edge_mat = empty
        g._edges_released = True


def to_networkx_route(graph):
    """Rebuild `edge_mat` through the literal networkx route of `util.py:43-103`.

    Used by tests to pin that the vectorised edge list and the reference's own
    construction coalesce to the same adjacency. Needs networkx.
    """
    import networkx as nx

    n = len(graph.g)
    half = graph.edge_mat.shape[1] // 2
    em = graph.edge_mat.numpy()
    conn = {}
    for i, j in zip(em[0, :half].tolist(), em[1, :half].tolist()):
        conn.setdefault(i, []).append(j)
    g = nx.Graph()
    for j in range(n):
        g.add_node(j)
        for k in conn.get(j, []):
            g.add_edge(j, k)
    edges = [list(pair) for pair in g.edges()]
    edges.extend([[i, j] for j, i in edges])
    neighbors = [[] for _ in range(n)]
    for i, j in g.edges():
        neighbors[i].append(j)
        neighbors[j].append(i)
    out = SynthGraph(n, graph.label, torch.LongTensor(edges).transpose(0, 1), graph.node_features)
    out.g = g
    out.neighbors = neighbors
    out.max_neighbor = max(len(x) for x in neighbors)
    return out
