"""GPU parity tests of each libgnm entry point (called through the C ABI via ops.py) against
(a) the numpy CSR oracle and (b) the contract-level CPU stand-in (tests/emul_ops.py, fp64
inside). Integer outputs must match bit-exactly; fp32 outputs within a scaled tolerance."""
import numpy as np
import pytest
import torch

import emul_ops
from helpers import assert_close
from oracle import csr_oracle
from graph_neural_mapping_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5      # fp32 kernels vs fp64-accumulating stand-in, error scaled by the tensor's max-abs


def rand_graph_edges(rng, n, p, dup=0):
    a = np.triu(rng.random((n, n)) < p, 1)
    i, j = np.nonzero(a)
    src, dst = np.concatenate([i, j]), np.concatenate([j, i])
    if dup and src.size:
        k = rng.integers(0, src.size, size=dup)
        src, dst = np.concatenate([src, src[k]]), np.concatenate([dst, dst[k]])
    perm = rng.permutation(src.size)
    return np.stack([src[perm], dst[perm]], 0).astype(np.int64)


def build_inputs(edge_mats, counts):
    edges = np.concatenate(edge_mats, 1) if edge_mats else np.zeros((2, 0), np.int64)
    eo = np.cumsum([0] + [e.shape[1] for e in edge_mats]).astype(np.int64)
    no = np.cumsum([0] + list(counts)).astype(np.int32)
    return (torch.from_numpy(edges).to(DEV), torch.from_numpy(eo).to(DEV), torch.from_numpy(no).to(DEV))


@pytest.mark.parametrize("self_loops", [False, True])
@pytest.mark.parametrize("case", ["ragged", "dups", "empty_graph", "single", "n400"])
def test_csr_build_bit_exact(case, self_loops):
    rng = np.random.default_rng(7)
    if case == "ragged":
        counts = [5, 33, 1, 64, 17, 100]
        ems = [rand_graph_edges(rng, n, 0.3) for n in counts]
    elif case == "dups":
        counts = [20, 31]
        ems = [rand_graph_edges(rng, n, 0.4, dup=25) for n in counts]
    elif case == "empty_graph":
        counts = [6, 9, 4]
        ems = [rand_graph_edges(rng, 6, 0.5), np.zeros((2, 0), np.int64), rand_graph_edges(rng, 4, 0.9)]
    elif case == "single":
        counts = [50]
        ems = [rand_graph_edges(rng, 50, 0.3)]
    else:
        gs = synth.make_graphs(3, 400, 30, 300, seed0=5)
        counts = [400] * 3
        ems = [g.edge_mat.numpy() for g in gs]
    e, eo, no = build_inputs(ems, counts)
    m = int(sum(counts))
    rp, ci, st = ops.csr_build(e, eo, no, len(counts), max(counts), m, self_loops, False)
    assert int(st.item()) == 0
    rp_o, ci_o = csr_oracle.multiset_csr(ems, counts, learn_eps=not self_loops)
    assert np.array_equal(rp.cpu().numpy().astype(np.int64), rp_o)
    assert np.array_equal(ci.cpu().numpy().astype(np.int64), ci_o)
    # same indices as the reference's own coalesced sparse tensor (graphcnn.py:84-106 + spmm)
    idx, _ = csr_oracle.block_diag_coo(ems, counts, learn_eps=not self_loops)
    ref = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.ones(idx.shape[1]), (m, m)).coalesce()
    dense = torch.zeros(m, m)
    rows = np.repeat(np.arange(m), np.diff(rp_o))
    dense.index_put_((torch.from_numpy(rows), torch.from_numpy(ci_o)), torch.ones(len(ci_o)), accumulate=True)
    assert torch.equal(dense, ref.to_dense())
    if case != "dups":
        refcsr = ref.to_sparse_csr()
        assert np.array_equal(refcsr.crow_indices().numpy(), rp_o) and np.array_equal(refcsr.col_indices().numpy(), ci_o)
    # local-column form + batch gather reproduces the global form for any slot order
    rpl, cil, _ = ops.csr_build(e, eo, no, len(counts), max(counts), m, self_loops, True)
    order = list(rng.permutation(len(counts)))
    nnz_g = [ems[i].shape[1] + (counts[i] if self_loops else 0) for i in range(len(counts))]
    nnz_base = np.cumsum([0] + nnz_g)
    no_h = np.cumsum([0] + list(counts))
    rp_addr = torch.tensor([rpl.data_ptr() + 4 * int(no_h[i]) for i in order], dtype=torch.int64, device=DEV)
    ci_addr = torch.tensor([cil.data_ptr() + 4 * int(nnz_base[i]) for i in order], dtype=torch.int64, device=DEV)
    no2 = torch.from_numpy(np.cumsum([0] + [counts[i] for i in order]).astype(np.int32)).to(DEV)
    zo2 = torch.from_numpy(np.cumsum([0] + [nnz_g[i] for i in order]).astype(np.int64)).to(DEV)
    rp2, ci2, _ = ops.csr_batch_gather(rp_addr, ci_addr, None, no2, zo2, len(order), m, int(sum(nnz_g)))
    rp_o2, ci_o2 = csr_oracle.multiset_csr([ems[i] for i in order], [counts[i] for i in order], learn_eps=not self_loops)
    assert np.array_equal(rp2.cpu().numpy().astype(np.int64), rp_o2)
    assert np.array_equal(ci2.cpu().numpy().astype(np.int64), ci_o2)


def test_csr_build_flags_bad_index():
    em = np.array([[0, 1, 7], [1, 0, 2]], dtype=np.int64)
    e, eo, no = build_inputs([em], [3])
    _, _, st = ops.csr_build(e, eo, no, 1, 3, 3, False, False)
    assert int(st.item()) == 1


def _structure(rng, counts, p, self_loops):
    ems = [rand_graph_edges(rng, n, p) for n in counts]
    e, eo, no = build_inputs(ems, counts)
    m = int(sum(counts))
    rp, ci, _ = ops.csr_build(e, eo, no, len(counts), max(counts), m, self_loops, False)
    return rp, ci, no, m


@pytest.mark.parametrize("f", [1, 3, 8, 12, 64, 100, 128, 400])
@pytest.mark.parametrize("mode,use_eps,use_map", [(0, True, False), (0, False, False), (1, True, False),
                                                   (2, True, False), (0, True, True), (1, False, True)])
def test_aggregate(f, mode, use_eps, use_map):
    rng = np.random.default_rng(f * 10 + mode)
    counts = [37, 64, 5, 90]
    rp, ci, no, m = _structure(rng, counts, 0.3, not use_eps)
    if mode != 1:
        # avoid 0/0 in the stand-in's transpose path; isolated nodes are covered in test_aggregate_isolated_nan
        pass
    torch.manual_seed(f)
    eps = torch.tensor([0.37], device=DEV) if use_eps else None
    bias = torch.randn(f, device=DEV) if use_map else None
    if use_map:
        table = torch.randn(23, f, device=DEV)
        smap = torch.randint(0, 23, (m,), dtype=torch.int32, device=DEV)
        src = table
    else:
        src, smap = torch.randn(m, f, device=DEV), None
    out = torch.empty(m, f, device=DEV)
    ops.aggregate(rp, ci, src, smap, out, mode, eps, bias)
    ref = torch.empty(m, f)
    emul_ops.aggregate(rp.cpu(), ci.cpu(), src.cpu(), smap.cpu() if use_map else None, ref, mode,
                       eps.cpu() if use_eps else None, bias.cpu() if use_map else None)
    if mode == 2:
        deg = (rp[1:] - rp[:-1]).cpu()
        keep = torch.ones(m, dtype=torch.bool)
        # rows whose neighbours all have deg > 0 (always true for symmetric graphs)
        assert_close(out.cpu()[keep], ref[keep], TOL, "aggregate")
    else:
        assert_close(out, ref, TOL, "aggregate")


def test_aggregate_isolated_nan_and_strided():
    em = np.array([[0, 1], [1, 0]], dtype=np.int64)        # node 2 isolated
    e, eo, no = build_inputs([em], [3])
    rp, ci, _ = ops.csr_build(e, eo, no, 1, 3, 3, False, False)
    src = torch.randn(3, 8, device=DEV)
    out = torch.empty(3, 8, device=DEV)
    ops.aggregate(rp, ci, src, None, out, 1, torch.tensor([0.0], device=DEV), None)   # graphcnn.py:157-158: 0/0
    assert torch.isnan(out[2]).all() and not torch.isnan(out[:2]).any()
    # strided views (a layer slice of a wider buffer)
    big_s, big_d = torch.randn(3, 24, device=DEV), torch.zeros(3, 40, device=DEV)
    ops.aggregate(rp, ci, big_s[:, 8:16], None, big_d[:, 16:24], 0, None, None)
    assert torch.allclose(big_d[0, 16:24], big_s[1, 8:16]) and torch.all(big_d[:, :16] == 0) and torch.all(big_d[:, 24:] == 0)


@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (37, 10, 8), (300, 64, 64), (513, 400, 64), (257, 64, 400), (1000, 128, 128),
                                   (129, 12, 12)])
@pytest.mark.parametrize("kn,pro,stats", [(False, False, True), (False, True, True), (True, False, False)])
def test_linear(m, k, n, kn, pro, stats):
    torch.manual_seed(m + k + n)
    x = torch.randn(m, k, device=DEV)
    w = torch.randn((k, n) if kn else (n, k), device=DEV) * 0.3
    b = torch.randn(n, device=DEV)
    sc = torch.rand(k, device=DEV) + 0.5 if pro else None
    sh = torch.randn(k, device=DEV) if pro else None
    y = torch.empty(m, n, device=DEV)
    st = torch.zeros(2 * n, dtype=torch.float64, device=DEV) if stats else None
    ops.linear(x, w, kn, b, sc, sh, y, st)
    yr = torch.empty(m, n)
    sr = torch.zeros(2 * n, dtype=torch.float64) if stats else None
    emul_ops.linear(x.cpu(), w.cpu(), kn, b.cpu(), sc.cpu() if pro else None, sh.cpu() if pro else None, yr, sr)
    assert_close(y, yr, TOL, "linear")
    if stats:
        assert_close(st, sr, 1e-5, "linear col_stats")
        st2 = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
        ops.col_stats(y, st2)
        assert_close(st2, sr, 1e-5, "col_stats")


@pytest.mark.parametrize("m,no,ni", [(1, 1, 1), (300, 8, 10), (5000, 64, 64), (1111, 64, 400), (2049, 128, 128), (700, 12, 0)])
@pytest.mark.parametrize("pro", [False, True])
def test_linear_wgrad(m, no, ni, pro):
    torch.manual_seed(m)
    dz = torch.randn(m, no, device=DEV)
    x = torch.randn(m, ni, device=DEV) if ni > 0 else None
    sc = torch.rand(ni, device=DEV) + 0.5 if (pro and ni > 0) else None
    sh = torch.randn(ni, device=DEV) if (pro and ni > 0) else None
    dw = torch.zeros(no, ni, device=DEV) if ni > 0 else None
    db = torch.zeros(no, device=DEV)
    ops.linear_wgrad(dz, x, sc, sh, dw, db)
    dwr = torch.zeros(no, ni) if ni > 0 else None
    dbr = torch.zeros(no)
    emul_ops.linear_wgrad(dz.cpu(), x.cpu() if x is not None else None, sc.cpu() if sc is not None else None,
                          sh.cpu() if sh is not None else None, dwr, dbr)
    if ni > 0:
        assert_close(dw, dwr, TOL, "dw")
    assert_close(db, dbr, TOL, "db")


@pytest.mark.parametrize("f", [8, 12, 64, 128, 300])
def test_batchnorm_pieces(f):
    torch.manual_seed(f)
    counts = [40, 7, 129]
    m = sum(counts)
    no = torch.tensor(np.cumsum([0] + counts).astype(np.int32), device=DEV)
    z = torch.randn(m, f, device=DEV) * 3 + 1
    gamma, beta = torch.rand(f, device=DEV) + 0.5, torch.randn(f, device=DEV)
    rm, rv = torch.randn(f, device=DEV), torch.rand(f, device=DEV) + 0.5
    nbt = torch.tensor(3, dtype=torch.int64, device=DEV)
    st = torch.zeros(2 * f, dtype=torch.float64, device=DEV)
    ops.col_stats(z, st)
    bufs = torch.empty(4, f, device=DEV)
    rm_c, rv_c, nbt_c = rm.cpu().clone(), rv.cpu().clone(), nbt.cpu().clone()
    ops.bn_finalize(st, float(m), gamma, beta, 1e-5, 0.1, rm, rv, nbt, bufs[0], bufs[1], bufs[2], bufs[3])
    # against torch's own batch_norm (what the reference calls)
    yr = torch.nn.functional.batch_norm(z.cpu(), rm_c, rv_c, gamma.cpu(), beta.cpu(), True, 0.1, 1e-5)
    assert_close(z * bufs[0] + bufs[1], yr, TOL, "bn train affine")
    assert_close(rm, rm_c, TOL, "running_mean")
    assert_close(rv, rv_c, TOL, "running_var")
    assert int(nbt) == int(nbt_c) + 1
    ev = torch.empty(4, f, device=DEV)
    ops.bn_eval_affine(rm, rv, gamma, beta, 1e-5, ev[0], ev[1], ev[2], ev[3])
    ye = torch.nn.functional.batch_norm(z.cpu(), rm_c, rv_c, gamma.cpu(), beta.cpu(), False, 0.1, 1e-5)
    assert_close(z * ev[0] + ev[1], ye, TOL, "bn eval affine")
    # apply + readout
    for ps in (None, torch.tensor([1.0 / c for c in counts], device=DEV)):
        h = torch.empty(m, f, device=DEV)
        wide = torch.zeros(3, 2 * f, device=DEV)
        ops.bn_relu_readout(z, bufs[0], bufs[1], h, no, 3, ps, wide[:, f:])
        hr, pr = torch.empty(m, f), torch.zeros(3, f)
        emul_ops.bn_relu_readout(z.cpu(), bufs[0].cpu(), bufs[1].cpu(), hr, no.cpu(), 3, ps.cpu() if ps is not None else None, pr)
        assert_close(h, hr, TOL, "h")
        assert_close(wide[:, f:], pr, TOL, "pooled")
        assert torch.all(wide[:, :f] == 0)
    # backward pieces with every gradient source switched on
    d_out = torch.randn(m, f, device=DEV)
    d_pool = torch.randn(3, 2 * f, device=DEV)
    d_score = torch.randn(m, device=DEV)
    u = torch.randn(3, 3 * f, device=DEV)
    d_neg = torch.randn(3, f, device=DEV)
    ps = torch.tensor([1.0 / c for c in counts], device=DEV)
    dy = torch.empty(m, f, device=DEV)
    bst = torch.zeros(2 * f, dtype=torch.float64, device=DEV)
    ops.relu_bn_bwd_reduce(z, bufs[0], bufs[1], bufs[2], bufs[3], d_out, d_pool[:, f:], ps, d_score, u[:, f:2 * f], d_neg, 3,
                           no, 3, dy, bst)
    dyr, bstr = torch.empty(m, f), torch.zeros(2 * f, dtype=torch.float64)
    c = lambda t: t.cpu() if t is not None else None
    emul_ops.relu_bn_bwd_reduce(c(z), c(bufs[0]), c(bufs[1]), c(bufs[2]), c(bufs[3]), c(d_out), c(d_pool)[:, f:], c(ps), c(d_score),
                                c(u)[:, f:2 * f], c(d_neg), 3, c(no), 3, dyr, bstr)
    assert_close(dy, dyr, TOL, "dy")
    assert_close(bst, bstr, 1e-5, "bwd stats")
    for use_stats in (True, False):
        d1, d2 = dy.clone(), dyr.clone()
        ops.bn_bwd_apply(z, bufs[2], bufs[3], gamma, bst if use_stats else None, float(m), d1)
        emul_ops.bn_bwd_apply(c(z), c(bufs[2]), c(bufs[3]), c(gamma), bstr if use_stats else None, float(m), d2)
        assert_close(d1, d2, TOL, "dz")
    # full check of the two-pass BatchNorm backward against autograd of torch's batch_norm + relu
    zz = z.cpu().clone().requires_grad_(True)
    g_, b_ = gamma.cpu().clone().requires_grad_(True), beta.cpu().clone().requires_grad_(True)
    out = torch.relu(torch.nn.functional.batch_norm(zz, None, None, g_, b_, True, 0.1, 1e-5))
    gz, gg, gb = torch.autograd.grad(out, [zz, g_, b_], d_out.cpu())
    dy2 = torch.empty(m, f, device=DEV)
    st2 = torch.zeros(2 * f, dtype=torch.float64, device=DEV)
    ops.relu_bn_bwd_reduce(z, bufs[0], bufs[1], bufs[2], bufs[3], d_out, None, None, None, None, None, 0, no, 3, dy2, st2)
    ops.bn_bwd_apply(z, bufs[2], bufs[3], gamma, st2, float(m), dy2)
    assert_close(dy2, gz, 1e-4, "bn+relu dz vs autograd")
    assert_close(st2[:f], gb, 1e-4, "dbeta")
    assert_close(st2[f:], gg, 1e-4, "dgamma")


@pytest.mark.parametrize("L,f,n,b", [(3, 8, 12, 4), (5, 64, 48, 6), (2, 12, 16, 5), (5, 64, 400, 3), (3, 64, 37, 5), (1, 64, 5, 2),
                                     (4, 64, 129, 2), (2, 64, 1, 3)])
def test_dgi_scores(L, f, n, b):
    torch.manual_seed(L * f)
    m = n * b
    h_all = torch.randn(L, m, f, device=DEV)
    u = torch.randn(b, L * f, device=DEV)
    bias = torch.tensor([0.3], device=DEV)
    perm = torch.randperm(b).to(torch.int32).to(DEV)
    no = torch.arange(0, m + 1, n, dtype=torch.int32, device=DEV)
    table = ops.gather_nf_rows(h_all, b)
    tr = emul_ops.gather_nf_rows(h_all.cpu(), b)
    assert torch.equal(table.cpu(), tr)
    out = torch.empty(2 * m, 1, device=DEV)
    ops.dgi_score_fwd(h_all, u, table, perm, no, b, bias, out)
    outr = torch.empty(2 * m, 1)
    emul_ops.dgi_score_fwd(h_all.cpu(), u.cpu(), tr, perm.cpu(), no.cpu(), b, bias.cpu(), outr)
    assert_close(out, outr, TOL, "dgi fwd")
    # literal reference formulation: nn.Bilinear on expanded c and n_f[idx] (graphcnn.py:198-201,241-246)
    w = torch.randn(1, L * f, L * f) * 0.1
    cvec = torch.rand(b, L * f)
    uu = (cvec @ w[0].t()).to(DEV)
    ops.dgi_score_fwd(h_all, uu, table, perm, no, b, bias, out)
    nf = torch.cat([h_all[l].cpu() for l in range(L)], 1)
    idx = np.repeat(perm.cpu().numpy(), n)
    cx = cvec.repeat_interleave(n, 0)
    lit = torch.cat([torch.nn.functional.bilinear(nf, cx, w, bias.cpu()),
                     torch.nn.functional.bilinear(nf[idx], cx, w, bias.cpu())], 0)
    assert_close(out, lit, 5e-5, "dgi fwd vs nn.Bilinear")
    d = torch.randn(2 * m, device=DEV)
    du, s2 = torch.empty(b, L * f, device=DEV), torch.empty(b, device=DEV)
    db = torch.zeros(1, dtype=torch.float64, device=DEV)
    ops.dgi_score_bwd(h_all, d, table, perm, no, b, du, s2, db)
    dur, s2r, dbr = torch.empty(b, L * f), torch.empty(b), torch.zeros(1, dtype=torch.float64)
    emul_ops.dgi_score_bwd(h_all.cpu(), d.cpu(), tr, perm.cpu(), no.cpu(), b, dur, s2r, dbr)
    assert_close(du, dur, TOL, "du")
    assert_close(s2, s2r, TOL, "s2")
    assert_close(db, dbr, 1e-5, "dbias")
    # standalone row-dot
    o2 = torch.empty(m, device=DEV)
    ops.rowdot_score(table.new_empty(0, 1) if False else torch.cat([h_all[l] for l in range(L)], 1).contiguous(), uu, n, bias, None, o2)
    assert_close(o2, lit[:m, 0], 5e-5, "rowdot")


def test_dot_rows_and_scatter():
    torch.manual_seed(0)
    m, f, t = 1234, 64, 50
    a, b = torch.randn(m, f, device=DEV), torch.randn(t, f, device=DEV)
    mp = torch.randint(0, t, (m,), dtype=torch.int32, device=DEV)
    out = torch.zeros(1, dtype=torch.float64, device=DEV)
    ops.dot_rows(a, b, mp, out)
    prod = a.double() * b[mp.long()].double()
    # cancelling sum: compare on the scale of sum |a*b| (fp32 products, per-row fp32 partials, fp64 across rows)
    assert abs(float(out) - float(prod.sum())) <= 1e-6 * float(prod.abs().sum()), "dot_rows"
    out.zero_()
    c = torch.randn(m, f, device=DEV)
    ops.dot_rows(a, c, None, out)
    prod = a.double() * c.double()
    assert abs(float(out) - float(prod.sum())) <= 1e-6 * float(prod.abs().sum()), "dot_rows nomap"
    for tt, ff in [(50, 64), (400, 64), (1000, 128), (7, 12)]:
        g = torch.randn(m, ff, device=DEV)
        tags = torch.randint(0, tt, (m,), dtype=torch.int32, device=DEV)
        tg = torch.zeros(tt, ff, device=DEV)
        ops.scatter_rows_add(g, tags, tg)
        ref = torch.zeros(tt, ff, dtype=torch.float64).index_add_(0, tags.cpu().long(), g.cpu().double())
        assert_close(tg, ref, TOL, "scatter_rows_add")


def test_no_cpu_fallback():
    x = torch.randn(4, 4)
    with pytest.raises(RuntimeError):
        ops.col_stats(x, torch.zeros(8, dtype=torch.float64))
    from graph_neural_mapping_b200.models import GIN_InfoMaxReg
    m = GIN_InfoMaxReg(2, 2, 5, 4, 2, 0.0, True, "sum", "sum", torch.device("cpu"))
    g = synth.make_graph(0, 5, 30, 32)
    with pytest.raises(RuntimeError):
        m([g])


def _bitmaps(rp_local, ci_local, no, counts):
    words = [n * ((n + 31) // 32) for n in counts]
    bo = torch.from_numpy(np.cumsum([0] + words).astype(np.int64)).to(DEV)
    bm, dup = ops.bitmap_build(rp_local, ci_local, no, bo, len(counts), int(sum(words)))
    addr = torch.tensor([bm.data_ptr() + 4 * int(o) for o in np.cumsum([0] + words)[:-1]], dtype=torch.int64, device=DEV)
    return bm, dup, addr, bo


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("counts", [[12, 12, 12], [400, 400], [37, 64, 5, 90, 1], [1000], [449, 130], [416, 129, 128, 400] * 40])
@pytest.mark.parametrize("f", [8, 12, 64, 128])
@pytest.mark.parametrize("mode,use_eps,use_map", [(0, True, False), (0, False, False), (1, True, False), (2, False, False),
                                                   (0, True, True)])
def test_aggregate_dense_matches_csr_and_oracle(counts, f, mode, use_eps, use_map, impl):
    if impl == 2 and max(counts) > 416 and mode == 1:
        pytest.skip("graphs above 416 nodes run the tcgen05 kernel in K-split passes; the forward average (a division of "
                    "the accumulated total) stays on the mma.sync kernel")
    if len(counts) > 100 and (f != 64 or mode != 0):
        pytest.skip("the many-graph case (persistent CTAs wrap around) is run for F=64, sum pooling")
    rng = np.random.default_rng(len(counts) * 100 + f + mode)
    self_loops = not use_eps
    ems = [rand_graph_edges(rng, n, 0.3) for n in counts]
    if mode != 0:
        # make sure no node is isolated (the host keeps such batches on the CSR kernel)
        ems = [np.concatenate([e, np.stack([np.arange(n), (np.arange(n) + 1) % n]), np.stack([(np.arange(n) + 1) % n, np.arange(n)])], 1)
               if n > 2 else e for e, n in zip(ems, counts)]
        ems = [np.unique(e, axis=1) for e in ems]
        ems = [e[:, e[0] != e[1]] for e in ems]
    e, eo, no = build_inputs(ems, counts)
    m = int(sum(counts))
    rp, ci, _ = ops.csr_build(e, eo, no, len(counts), max(counts), m, self_loops, False)
    rpl, cil, _ = ops.csr_build(e, eo, no, len(counts), max(counts), m, self_loops, True)
    bm, dup, addr, _ = _bitmaps(rpl, cil, no, counts)
    assert int(dup.sum()) == 0
    if mode != 0 and bool(((rp[1:] - rp[:-1]) == 0).any()):
        pytest.skip("isolated node in an average-pooling case")
    torch.manual_seed(f)
    eps = torch.tensor([-0.21], device=DEV) if use_eps else None
    bias = torch.randn(f, device=DEV) if use_map else None
    if use_map:
        src = torch.randn(29, f, device=DEV) * 5
        smap = torch.randint(0, 29, (m,), dtype=torch.int32, device=DEV)
    else:
        src, smap = torch.randn(m, f, device=DEV) * 5 + 1, None
    a = torch.full((m, f), float("nan"), device=DEV)
    b = torch.empty(m, f, device=DEV)
    ops.aggregate_dense(addr, no, rp, len(counts), max(counts), src, smap, a, mode, eps, bias, impl=impl)
    assert not ops.aggregate_tc_status(), "tcgen05 kernel hit a barrier timeout"
    ops.aggregate(rp, ci, src, smap, b, mode, eps, bias)
    if len(counts) > 100:
        assert_close(a, b, 1e-5, "dense vs csr (many graphs)")
        return
    ref = torch.empty(m, f)
    emul_ops.aggregate(rp.cpu(), ci.cpu(), src.cpu(), smap.cpu() if use_map else None, ref, mode,
                       eps.cpu() if use_eps else None, bias.cpu() if use_map else None)
    assert_close(a, ref, TOL, "dense vs fp64")
    assert_close(b, ref, TOL, "csr vs fp64")
    # the bf16x3 split is exact, so the two kernels differ only by fp32 summation order
    assert_close(a, b, 1e-5, "dense vs csr")


def test_bitmap_build_bits_and_duplicates():
    rng = np.random.default_rng(3)
    counts = [33, 64, 7]
    ems = [rand_graph_edges(rng, 33, 0.4), rand_graph_edges(rng, 64, 0.2, dup=5), rand_graph_edges(rng, 7, 0.9)]
    e, eo, no = build_inputs(ems, counts)
    rpl, cil, _ = ops.csr_build(e, eo, no, 3, 64, sum(counts), True, True)
    bm, dup, addr, bo = _bitmaps(rpl, cil, no, counts)
    assert dup.cpu().tolist() == [0, 1, 0]
    bm_h = bm.cpu().numpy().view(np.uint32)
    bo_h = bo.cpu().numpy()
    for g, n in enumerate(counts):
        w = (n + 31) // 32
        words = bm_h[bo_h[g]:bo_h[g] + n * w].reshape(n, w)
        bits = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(n, w * 32)
        dense = (csr_oracle.dense_adjacency([ems[g]], [n], learn_eps=False) > 0).astype(np.uint32)
        assert np.array_equal(bits[:, :n], dense) and not bits[:, n:].any()


def test_aggregate_dense_exactness_on_wide_dynamic_range():
    """hi+mid+lo reconstructs fp32 exactly: with a single neighbour per row the dense kernel must return the
    source row bit-for-bit, across 60 binades of magnitude."""
    n = 64
    em = np.stack([np.arange(n), (np.arange(n) + 1) % n]).astype(np.int64)     # row i has the single neighbour i+1
    e, eo, no = build_inputs([em], [n])
    rp, ci, _ = ops.csr_build(e, eo, no, 1, n, n, False, False)
    rpl, cil, _ = ops.csr_build(e, eo, no, 1, n, n, False, True)
    bm, dup, addr, _ = _bitmaps(rpl, cil, no, [n])
    torch.manual_seed(1)
    src = torch.randn(n, 64, device=DEV) * torch.exp2(torch.randint(-30, 30, (n, 64), device=DEV).float())
    for impl in (1, 2):
        out = torch.empty(n, 64, device=DEV)
        ops.aggregate_dense(addr, no, rp, 1, n, src, None, out, 0, None, None, impl=impl)
        assert torch.equal(out, src.roll(-1, 0)), "impl %d" % impl
    assert not ops.aggregate_tc_status()


@pytest.mark.parametrize("counts,f,mode,eps_on,terms", [([12, 12, 12], 8, 0, True, "all"), ([400, 400, 400], 64, 0, False, "all"),
                                                        ([37, 64, 5, 90, 1], 64, 0, True, "none"),
                                                        ([400, 129, 128, 399] * 40, 64, 0, False, "all"), ([416, 100], 64, 0, False, "all"),
                                                        ([48, 48], 12, 2, False, "pooled"), ([400] * 5, 128, 0, False, "all")])
def test_aggregate_dense_relu_bn_bwd(counts, f, mode, eps_on, terms):
    """The backward aggregation fused with relu / BatchNorm backward of the layer below must equal aggregate_dense
    followed by relu_bn_bwd_reduce: dy, and the reduction [sum dy, sum dy*xhat]; wider layers return False untouched."""
    rng = np.random.default_rng(len(counts) * 11 + f)
    ems = [rand_graph_edges(rng, n, 0.3) for n in counts]
    if mode != 0:
        ems = [np.concatenate([e, np.stack([np.arange(n), (np.arange(n) + 1) % n]), np.stack([(np.arange(n) + 1) % n, np.arange(n)])], 1)
               for e, n in zip(ems, counts)]
        ems = [np.unique(e, axis=1) for e in ems]
        ems = [e[:, e[0] != e[1]] for e in ems]
    e, eo, no = build_inputs(ems, counts)
    m, b = int(sum(counts)), len(counts)
    rp, ci, _ = ops.csr_build(e, eo, no, b, max(counts), m, not eps_on, False)
    rpl, cil, _ = ops.csr_build(e, eo, no, b, max(counts), m, not eps_on, True)
    bm, dup, addr, _ = _bitmaps(rpl, cil, no, counts)
    torch.manual_seed(f + m)
    src, z = torch.randn(m, f, device=DEV), torch.randn(m, f, device=DEV)
    scale, shift = torch.rand(f, device=DEV) + 0.5, torch.randn(f, device=DEV) * 0.3
    mean, rstd = torch.randn(f, device=DEV), torch.rand(f, device=DEV) + 0.5
    eps = torch.tensor([0.3], device=DEV) if eps_on else None
    L = 3
    d_pooled = torch.randn(b, L * f, device=DEV)[:, f:2 * f] if terms in ("all", "pooled") else None      # strided slice
    pool_scale = torch.rand(b, device=DEV) + 0.5 if terms == "all" else None
    d_score = torch.randn(m, device=DEV) if terms == "all" else None
    u = torch.randn(b, L * f, device=DEV)[:, f:2 * f] if terms == "all" else None
    n_neg = min(m, 7 * b) if terms == "all" else 0
    d_neg = torch.randn(max(n_neg, 1), L * f, device=DEV)[:, f:2 * f] if terms == "all" else None
    dy = torch.full((m, f), float("nan"), device=DEV)
    st = torch.zeros(2 * f, dtype=torch.float64, device=DEV)
    launched = ops.aggregate_dense_relu_bn_bwd(addr, no, rp, b, max(counts), src, mode, eps, z, scale, shift, mean, rstd,
                                               d_pooled, pool_scale, d_score, u, d_neg, n_neg, dy, st)
    if f > 64 or max(counts) > 400:
        # wider layers / graphs above 400 nodes (no room for the z tiles next to the operand planes): untouched
        assert launched is False and bool(torch.isnan(dy).all()) and float(st.abs().sum()) == 0.0
        return
    assert launched is True
    assert not ops.aggregate_tc_status(), "tcgen05 kernel hit a barrier timeout"
    d_h = torch.empty(m, f, device=DEV)
    ops.aggregate_dense(addr, no, rp, b, max(counts), src, None, d_h, mode, eps, None, impl=2)
    dy2 = torch.empty(m, f, device=DEV)
    st2 = torch.zeros(2 * f, dtype=torch.float64, device=DEV)
    ops.relu_bn_bwd_reduce(z, scale, shift, mean, rstd, d_h, d_pooled, pool_scale, d_score, u, d_neg, n_neg, no, b, dy2, st2)
    assert_close(dy, dy2, 1e-6, "fused dy vs aggregate + relu_bn_bwd_reduce")
    assert torch.equal(dy == 0, dy2 == 0), "same ReLU mask"
    assert_close(st, st2, 1e-6, "fused BatchNorm-backward reduction")


@pytest.mark.parametrize("counts,f,mode", [([12, 12, 12], 8, 0), ([400, 400, 400], 64, 0), ([37, 64, 5, 90, 1], 64, 0),
                                           ([400, 129, 128, 399] * 40, 64, 0), ([416, 129, 128, 400], 64, 0), ([48, 48], 12, 2),
                                           ([1000], 64, 0)])
def test_aggregate_dense_affine(counts, f, mode):
    """Aggregation of a BatchNorm-backward result folded into the row loads: Agg(cA*dy + cB*z + cC) must equal
    bn_bwd_apply followed by the aggregation, and the fp64 stand-in; batches the tcgen05 kernel cannot take return False
    without launching anything."""
    rng = np.random.default_rng(len(counts) * 7 + f)
    ems = [rand_graph_edges(rng, n, 0.3) for n in counts]
    if mode != 0:
        ems = [np.concatenate([e, np.stack([np.arange(n), (np.arange(n) + 1) % n]), np.stack([(np.arange(n) + 1) % n, np.arange(n)])], 1)
               for e, n in zip(ems, counts)]
        ems = [np.unique(e, axis=1) for e in ems]
        ems = [e[:, e[0] != e[1]] for e in ems]
    e, eo, no = build_inputs(ems, counts)
    m = int(sum(counts))
    self_loops = True                        # the folded path serves learn_eps == False models (self loop in Adj_block)
    rp, ci, _ = ops.csr_build(e, eo, no, len(counts), max(counts), m, self_loops, False)
    rpl, cil, _ = ops.csr_build(e, eo, no, len(counts), max(counts), m, self_loops, True)
    bm, dup, addr, _ = _bitmaps(rpl, cil, no, counts)
    torch.manual_seed(f + m)
    dy, z = torch.randn(m, f, device=DEV), torch.randn(m, f, device=DEV) * 2 + 0.3
    gamma, mean, rstd = torch.rand(f, device=DEV) + 0.5, torch.randn(f, device=DEV), torch.rand(f, device=DEV) + 0.5
    stats = torch.randn(2 * f, dtype=torch.float64, device=DEV) * m
    coef = torch.empty(3, f, device=DEV)
    ops.bn_bwd_coeffs(stats, float(m), gamma, mean, rstd, coef)
    out = torch.full((m, f), float("nan"), device=DEV)
    launched = ops.aggregate_dense_affine(addr, no, rp, len(counts), max(counts), dy, z, coef, out, mode)
    if max(counts) > 400:                 # the second stream's landing area leaves shared memory for 400-node planes
        assert launched is False and bool(torch.isnan(out).all())
        return
    assert launched is True
    assert not ops.aggregate_tc_status(), "tcgen05 kernel hit a barrier timeout"
    dz = dy.clone()
    ops.bn_bwd_apply(z, mean, rstd, gamma, stats, float(m), dz)
    two_pass = torch.empty(m, f, device=DEV)
    ops.aggregate_dense(addr, no, rp, len(counts), max(counts), dz, None, two_pass, mode, None, None, impl=2)
    assert_close(out, two_pass, 2e-5, "folded vs bn_bwd_apply + aggregate")
    if len(counts) <= 8:
        ref = torch.empty(m, f)
        dzr = (coef[0].double() * dy.double() + coef[1].double() * z.double() + coef[2].double()).cpu()
        emul_ops.aggregate(rp.cpu(), ci.cpu(), dzr.float(), None, ref, mode, None, None)
        assert_close(out, ref, 2e-5, "folded vs fp64")


@pytest.mark.parametrize("sizes,f,p,self_loops,eps", [([12, 12, 12], 8, 0.3, False, True), ([40, 17, 1, 33], 64, 0.2, True, False),
                                                      ([400, 400], 64, 0.3, False, True), ([9, 5], 70, 0.0, False, False),
                                                      ([30], 3, 0.5, True, False)])
def test_aggregate_max(sizes, f, p, self_loops, eps):
    """Max pooling over neighbours (graphcnn.py:137-143): values, arg-max routing of the gradient, the dummy row
    (column minimum) for nodes without any entry, the (1 + eps) self term; against the literal padded-list gather."""
    rng = np.random.default_rng(sum(sizes) + f)
    mats = [rand_graph_edges(rng, n, p) for n in sizes]
    e, eo, no = build_inputs(mats, sizes)
    m = sum(sizes)
    rp, ci, _ = ops.csr_build(e, eo, no, len(sizes), max(sizes), m, self_loops, False)
    torch.manual_seed(f)
    h = torch.randn(m, f, device=DEV)
    h[::3] = torch.relu(h[::3])                                  # exact zeros / ties like a post-ReLU layer
    ep = torch.tensor([0.37], device=DEV) if eps else None
    cmin = ops.col_min(h)
    out = torch.empty(m, f, device=DEV)
    amax = torch.empty(m, f, dtype=torch.int32, device=DEV)
    ops.aggregate_max(rp, ci, h, cmin, ep, out, amax)
    # literal reference: padded neighbour list, -1 pads -> dummy row = column minimum
    hc = h.cpu().double().requires_grad_()
    rpc, cic = rp.cpu().long(), ci.cpu().long()
    deg = (rpc[1:] - rpc[:-1])
    width = max(int(deg.max()), 1)
    padded = torch.full((m, width), -1, dtype=torch.long)
    for i in range(m):
        padded[i, :int(deg[i])] = cic[rpc[i]:rpc[i + 1]]
    dummy = hc.min(0)[0]
    hw = torch.cat([hc, dummy.reshape(1, -1)], 0)
    ref = hw[padded].max(1)[0]
    if eps:
        ref = ref + (1 + 0.37) * hc
    assert_close(out, ref.detach(), 1e-6, "max pooling forward")
    am = amax.cpu().long()
    assert bool(((am == m) == (deg == 0).unsqueeze(1)).all()), "dummy exactly for empty rows"
    d_out = torch.randn(m, f, device=DEV)
    d_h = torch.full((m, f), float("nan"), device=DEV)
    ops.aggregate_max_bwd(rp, ci, d_out, amax, cmin, ep, d_h)
    ref.backward(d_out.cpu().double())
    # ties (exact zeros) may be routed to another tied row than torch's choice: compare where the maximum is unique
    vals = hw.detach()[padded]
    n_at_max = (vals == vals.max(1, keepdim=True)[0]).sum(1)
    unique_rows = (n_at_max <= 1).all(1) | (deg == 0)
    if bool(unique_rows.all()):
        assert_close(d_h, hc.grad, 1e-5, "max pooling backward")
    # always: the gradient mass per column is conserved
    assert_close(d_h.double().sum(0), hc.grad.sum(0), 1e-5, "column sums of the routed gradient")
    d_h2 = torch.empty(m, f, device=DEV)
    ops.aggregate_max_bwd(rp, ci, d_out, amax, cmin, ep, d_h2)
    # pull form: bit-reproducible, except for the dummy hits of isolated nodes (fp32 atomics onto the column-minimum row)
    assert torch.equal(d_h, d_h2) or bool((deg == 0).any()), "deterministic without dummy hits"


@pytest.mark.parametrize("n_graphs,period,f,table", [(1, 5, 4, 5), (3, 7, 12, 9), (40, 400, 64, 400), (1024, 400, 64, 400),
                                                     (33, 50, 8, 64)])
def test_rows_period_sum(n_graphs, period, f, table):
    """Layer-0 table gradient for batches that share one injective tag sequence: equals the scatter, is
    deterministic, accumulates into the output, drops out-of-range tags."""
    torch.manual_seed(n_graphs + period)
    g = torch.randn(n_graphs * period, f, device=DEV)
    seq = torch.randperm(table, device=DEV)[:period].to(torch.int32)
    tags = seq.repeat(n_graphs)
    base = torch.randn(table, f, device=DEV)
    out = base.clone()
    ops.rows_period_sum(g, period, tags, out)
    ref = base.cpu().double().index_add_(0, tags.cpu().long(), g.cpu().double())
    assert_close(out, ref, TOL, "rows_period_sum vs scatter reference")
    out2 = base.clone()
    ops.rows_period_sum(g, period, tags, out2)
    assert torch.equal(out, out2), "deterministic"
    sc = base.clone()
    ops.scatter_rows_add(g, tags, sc)
    assert_close(out, sc, TOL, "rows_period_sum vs scatter_rows_add")
    if period <= table:
        ident = torch.zeros(table, f, device=DEV)
        ops.rows_period_sum(g, period, None, ident)
        ref_i = torch.zeros(table, f, dtype=torch.float64)
        ref_i[:period] = g.cpu().double().view(n_graphs, period, f).sum(0)
        assert_close(ident, ref_i, TOL, "identity tags")
    bad = tags.clone()
    bad[:period][0] = table + 3                      # out-of-range tag of position 0: that position is dropped
    out3 = torch.zeros(table, f, device=DEV)
    ops.rows_period_sum(g, period, bad, out3)
    ref3 = torch.zeros(table, f, dtype=torch.float64)
    s = g.cpu().double().view(n_graphs, period, f).sum(0)
    ref3.index_add_(0, seq.cpu().long()[1:], s[1:])
    assert_close(out3, ref3, TOL, "out-of-range tag dropped")
    with pytest.raises(Exception):
        ops.rows_period_sum(g[:-1], period, tags, out) if period > 1 else (_ for _ in ()).throw(RuntimeError("n/a"))


@pytest.mark.parametrize("m,fo,fi", [(1, 1, 1), (300, 8, 12), (4097, 64, 64), (1000, 12, 64), (640, 64, 8), (129, 63, 37)])
@pytest.mark.parametrize("act,train", [(True, True), (False, True), (True, False)])
def test_linear_bwd_fused(m, fo, fi, act, train):
    torch.manual_seed(m + fo)
    dy, z = torch.randn(m, fo, device=DEV), torch.randn(m, fo, device=DEV) * 2 + 0.5
    x = torch.randn(m, fi, device=DEV)
    w = torch.randn(fo, fi, device=DEV) * 0.3
    gamma, mean, rstd = torch.rand(fo, device=DEV) + 0.5, torch.randn(fo, device=DEV), torch.rand(fo, device=DEV) + 0.5
    stats = torch.randn(2 * fo, dtype=torch.float64, device=DEV) * m if train else None
    coef = torch.empty(3, fo, device=DEV)
    ops.bn_bwd_coeffs(stats, float(m), gamma, mean, rstd, coef)
    coef_r = torch.empty(3, fo)
    c = lambda t: t.cpu() if t is not None else None
    emul_ops.bn_bwd_coeffs(c(stats), float(m), c(gamma), c(mean), c(rstd), coef_r)
    assert_close(coef, coef_r, 1e-5, "coef")
    # the affine form equals the two-pass BatchNorm backward
    if train:
        dref = dy.clone()
        ops.bn_bwd_apply(z, mean, rstd, gamma, stats, float(m), dref)
        assert_close(coef[0] * dy + coef[1] * z + coef[2], dref, 2e-5, "dz affine form")
    isc = torch.rand(fi, device=DEV) + 0.5 if act else None
    ish = torch.randn(fi, device=DEV) if act else None
    imu = torch.randn(fi, device=DEV) if act else None
    irs = torch.rand(fi, device=DEV) + 0.5 if act else None
    dw, db = torch.zeros(fo, fi, device=DEV), torch.zeros(fo, device=DEV)
    dx = torch.full((m, fi), float("nan"), device=DEV)
    st = torch.zeros(2 * fi, dtype=torch.float64, device=DEV) if act else None
    ops.linear_bwd(dy, z, coef, x, isc, ish, imu, irs, w, dw, db, dx, st)
    dwr, dbr, dxr = torch.zeros(fo, fi), torch.zeros(fo), torch.empty(m, fi)
    str_ = torch.zeros(2 * fi, dtype=torch.float64) if act else None
    emul_ops.linear_bwd(c(dy), c(z), c(coef), c(x), c(isc), c(ish), c(imu), c(irs), c(w), dwr, dbr, dxr, str_)
    assert_close(dw, dwr, TOL, "dw")
    assert_close(db, dbr, TOL, "db")
    assert_close(dx, dxr, TOL, "dx")
    if act:
        assert_close(st, str_, 2e-5, "stats_in")
    # dx == None (first unit of layer 0 in training) still produces dw / db
    dw2, db2 = torch.zeros(fo, fi, device=DEV), torch.zeros(fo, device=DEV)
    ops.linear_bwd(dy, z, coef, x, None, None, None, None, w, dw2, db2, None, None)
    dwr2, dbr2 = torch.zeros(fo, fi), torch.zeros(fo)
    emul_ops.linear_bwd(c(dy), c(z), c(coef), c(x), None, None, None, None, c(w), dwr2, dbr2, None, None)
    assert_close(dw2, dwr2, TOL, "dw (no dx)")


@pytest.mark.parametrize("m,fo,fi", [(1, 1, 1), (300, 8, 12), (4097, 64, 64), (40000, 64, 64), (20001, 48, 64), (9000, 64, 33),
                                     (129, 63, 37), (64 * 148 * 3 + 5, 64, 64)])
@pytest.mark.parametrize("act,train", [(True, True), (False, True), (True, False)])
def test_linear_bwd_tcgen05(m, fo, fi, act, train):
    """Tensor-core backward unit - the one-pass kernel (impl 2; 64 x 64 aligned units with dx) and the two-pass pair of an
    input-gradient and a weight-gradient kernel (impl 3, and every other shape), bf16x3 exact splits - against the
    fp64 stand-in at the FFMA kernel's tolerance, and against the FFMA kernel itself."""
    torch.manual_seed(m + fo + fi)
    dy, z = torch.randn(m, fo, device=DEV), torch.randn(m, fo, device=DEV) * 2 + 0.5
    x = torch.randn(m, fi, device=DEV)
    w = torch.randn(fo, fi, device=DEV) * 0.3
    gamma, mean, rstd = torch.rand(fo, device=DEV) + 0.5, torch.randn(fo, device=DEV), torch.rand(fo, device=DEV) + 0.5
    stats = torch.randn(2 * fo, dtype=torch.float64, device=DEV) * m if train else None
    coef = torch.empty(3, fo, device=DEV)
    ops.bn_bwd_coeffs(stats, float(m), gamma, mean, rstd, coef)
    isc = torch.rand(fi, device=DEV) + 0.5 if act else None
    ish = torch.randn(fi, device=DEV) if act else None
    imu = torch.randn(fi, device=DEV) if act else None
    irs = torch.rand(fi, device=DEV) + 0.5 if act else None
    outs = []
    try:
        for impl in (2, 3, 1):
            ops.set_linear_impl(impl)
            before = ops.launch_counts()
            dw, db = torch.zeros(fo, fi, device=DEV), torch.zeros(fo, device=DEV)
            dx = torch.full((m, fi), float("nan"), device=DEV)
            st = torch.zeros(2 * fi, dtype=torch.float64, device=DEV) if act else None
            ops.linear_bwd(dy, z, coef, x, isc, ish, imu, irs, w, dw, db, dx, st)
            dw2, db2 = torch.zeros(fo, fi, device=DEV), torch.zeros(fo, device=DEV)
            ops.linear_bwd(dy, z, coef, x, None, None, None, None, w, dw2, db2, None, None)
            outs.append((dw, db, dx, st, dw2, db2))
            onepass = ops.launch_counts()["linear_bwd_onepass_tc"] - before["linear_bwd_onepass_tc"]
            assert onepass == (1 if impl == 2 and fo == 64 and fi == 64 else 0), (impl, onepass)
    finally:
        ops.set_linear_impl(0)
    assert not ops.aggregate_tc_status(), "tcgen05 kernel hit a barrier timeout"
    c = lambda t: t.cpu() if t is not None else None
    dwr, dbr, dxr = torch.zeros(fo, fi), torch.zeros(fo), torch.empty(m, fi)
    str_ = torch.zeros(2 * fi, dtype=torch.float64) if act else None
    emul_ops.linear_bwd(c(dy), c(z), c(coef), c(x), c(isc), c(ish), c(imu), c(irs), c(w), dwr, dbr, dxr, str_)
    dwr2, dbr2 = torch.zeros(fo, fi), torch.zeros(fo)
    emul_ops.linear_bwd(c(dy), c(z), c(coef), c(x), None, None, None, None, c(w), dwr2, dbr2, None, None)
    tc, tc2, ff = outs
    if act:
        # the ReLU mask is discontinuous: entries whose pre-activation is within rounding of zero may legitimately
        # flip between the fp32 fmaf and the fp64 stand-in; they are compared against the FFMA kernel only
        edge = (x.double() * isc.double() + ish.double()).abs() < 1e-5
        assert int(edge.sum()) < 1e-4 * edge.numel() + 16
        tc[2][edge] = 0.0
        tc2[2][edge] = 0.0
        dxr[edge.cpu()] = 0.0
        ff[2][edge] = 0.0
    assert_close(tc[2], dxr, TOL, "dx vs fp64")
    assert_close(tc[0], dwr, TOL, "dw vs fp64")
    assert_close(tc[1], dbr, TOL, "db vs fp64")
    assert_close(tc[4], dwr2, TOL, "dw (no dx, plain input) vs fp64")
    assert_close(tc[5], dbr2, TOL, "db (no dx) vs fp64")
    assert_close(tc[2], ff[2], TOL, "dx vs FFMA")
    assert_close(tc[0], ff[0], TOL, "dw vs FFMA")
    assert_close(tc2[2], dxr, TOL, "two-pass dx vs fp64")
    assert_close(tc2[0], dwr, TOL, "two-pass dw vs fp64")
    assert_close(tc2[1], dbr, TOL, "two-pass db vs fp64")
    if act:
        assert_close(tc[3], ff[3], 2e-5, "stats_in vs FFMA")
        assert_close(tc2[3], ff[3], 2e-5, "two-pass stats_in vs FFMA")
        if not bool(edge.any()):
            assert_close(tc[3], str_, 2e-5, "stats_in vs fp64")


@pytest.mark.parametrize("m", [1, 100, 128, 129, 4096 + 31, 128 * 148 + 1])
@pytest.mark.parametrize("act", [True, False])
def test_linear_bwd_onepass_edges(m, act):
    """The one-pass backward unit at its edges: fewer rows than one tile / fewer tiles than SMs / one tile more than SMs,
    every operand a column slice of a wider matrix (leading dimension 192), dw a slice of a wider gradient buffer that
    already holds values (the kernel ADDS), against the fp64 stand-in; untouched neighbours of the slices stay untouched."""
    fo = fi = 64
    torch.manual_seed(7 * m + act)
    wide = lambda rows: torch.randn(rows, 3 * 64, device=DEV)
    dy_w, z_w, x_w = wide(m), wide(m) * 2 + 0.5, wide(m)
    dy, z, x = dy_w[:, 64:128], z_w[:, 64:128], x_w[:, 64:128]
    w = torch.randn(fo, fi, device=DEV) * 0.3
    coef = torch.randn(3, fo, device=DEV)
    isc = torch.rand(fi, device=DEV) + 0.5 if act else None
    ish = torch.randn(fi, device=DEV) if act else None
    imu = torch.randn(fi, device=DEV) if act else None
    irs = torch.rand(fi, device=DEV) + 0.5 if act else None
    dx_w = torch.full((m, 3 * 64), 7.0, device=DEV)
    dw_w = torch.randn(fo, 2 * fi, device=DEV)
    dw0 = dw_w.clone()
    db = torch.randn(fo, device=DEV)
    db0 = db.clone()
    st = torch.zeros(2 * fi, dtype=torch.float64, device=DEV) if act else None
    try:
        ops.set_linear_impl(2)
        before = ops.launch_counts()["linear_bwd_onepass_tc"]
        ops.linear_bwd(dy, z, coef, x, isc, ish, imu, irs, w, dw_w[:, fi:], db, dx_w[:, 64:128], st)
        assert ops.launch_counts()["linear_bwd_onepass_tc"] == before + 1
    finally:
        ops.set_linear_impl(0)
    assert not ops.aggregate_tc_status(), "tcgen05 kernel hit a barrier timeout"
    c = lambda t: t.cpu().contiguous() if t is not None else None
    dwr, dbr, dxr = torch.zeros(fo, fi), torch.zeros(fo), torch.empty(m, fi)
    str_ = torch.zeros(2 * fi, dtype=torch.float64) if act else None
    emul_ops.linear_bwd(c(dy), c(z), c(coef), c(x), c(isc), c(ish), c(imu), c(irs), c(w), dwr, dbr, dxr, str_)
    dx = dx_w[:, 64:128].clone()
    if act:
        edge = (x.double() * isc.double() + ish.double()).abs() < 1e-5
        dx[edge] = 0.0
        dxr[edge.cpu()] = 0.0
    assert_close(dx, dxr, TOL, "dx")
    assert_close(dw_w[:, fi:] - dw0[:, fi:], dwr, TOL, "dw (added to the slice)")
    assert_close(db - db0, dbr, TOL, "db (added)")
    assert torch.equal(dw_w[:, :fi], dw0[:, :fi]) and bool((dx_w[:, :64] == 7.0).all()) and bool((dx_w[:, 128:] == 7.0).all())
    if act and not bool(edge.any()):
        assert_close(st, str_, 2e-5, "stats_in")


@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (37, 10, 8), (300, 64, 64), (5000, 64, 64), (129, 12, 12), (20000, 48, 64),
                                   (4097, 64, 33), (40000, 64, 64)])
@pytest.mark.parametrize("kn,pro,stats", [(False, False, True), (False, True, True), (True, False, False), (True, True, True)])
def test_linear_tcgen05(m, k, n, kn, pro, stats):
    """The tensor-core Linear (bf16x3 exact operand splits, six tcgen05 MMAs per k-step) against the fp64
    stand-in at the same tolerance as the fp32 FFMA kernel, and against the FFMA kernel itself."""
    torch.manual_seed(m + k + n)
    x = torch.randn(m, k, device=DEV) * 3
    w = torch.randn((k, n) if kn else (n, k), device=DEV) * 0.3
    b = torch.randn(n, device=DEV)
    sc = torch.rand(k, device=DEV) + 0.5 if pro else None
    sh = torch.randn(k, device=DEV) if pro else None
    try:
        outs = []
        for impl in (2, 1):
            ops.set_linear_impl(impl)
            y = torch.full((m, n), float("nan"), device=DEV)
            st = torch.zeros(2 * n, dtype=torch.float64, device=DEV) if stats else None
            ops.linear(x, w, kn, b, sc, sh, y, st)
            outs.append((y, st))
    finally:
        ops.set_linear_impl(0)
    assert not ops.aggregate_tc_status(), "tcgen05 kernel hit a barrier timeout"
    yr = torch.empty(m, n)
    sr = torch.zeros(2 * n, dtype=torch.float64) if stats else None
    emul_ops.linear(x.cpu(), w.cpu(), kn, b.cpu(), sc.cpu() if pro else None, sh.cpu() if pro else None, yr, sr)
    assert_close(outs[0][0], yr, TOL, "tcgen05 linear vs fp64")
    assert_close(outs[0][0], outs[1][0], TOL, "tcgen05 vs FFMA")
    if stats:
        assert_close(outs[0][1], sr, 1e-5, "tcgen05 col_stats")


def test_peer_exchange_protocol_two_ranks_on_one_gpu():
    """The NVLink peer-memory all-reduce (include/gnm.h, data-parallel section) with both "ranks" on this GPU: two
    exchange buffers, two communicators, the two kernels on two streams meet through the flag protocol. 100 exchanges
    of varying length back to back (both slot parities, counter in device memory), then the fused users: bn_finalize
    and bn_bwd_coeffs with a communicator must equal the same calls on pre-summed statistics."""
    import ctypes
    from graph_neural_mapping_b200 import lib as glib
    L = glib.load()
    world = 2
    bufs = []
    for _ in range(world):
        ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        glib.check(L.gnm_p2p_alloc(ctypes.byref(ptr), handle), "gnm_p2p_alloc")
        bufs.append(ptr)
    try:
        peers = torch.tensor([b.value for b in bufs], dtype=torch.int64, device=DEV)
        counters = [torch.zeros(1, dtype=torch.int32, device=DEV) for _ in range(world)]
        comms = [ops._P2PStruct(peers.data_ptr(), counters[r].data_ptr(), r, world) for r in range(world)]
        streams = [torch.cuda.Stream() for _ in range(world)]
        torch.manual_seed(3)
        torch.cuda.synchronize()
        for it in range(100):
            n = [1, 2, 128, 256, 37][it % 5]
            xs = [torch.randn(n, dtype=torch.float64, device=DEV) * 10.0 ** (it % 5 - 2) for _ in range(world)]
            want = xs[0] + xs[1]
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    glib.check(L.gnm_p2p_allreduce(ctypes.c_void_p(xs[r].data_ptr()), n, ctypes.byref(comms[r]),
                                                   ctypes.c_void_p(streams[r].cuda_stream)), "gnm_p2p_allreduce")
            torch.cuda.synchronize()
            assert torch.equal(xs[0], want) and torch.equal(xs[1], want), "exchange %d" % it
        aborted = ctypes.c_int(0)
        glib.check(L.gnm_p2p_status(ctypes.byref(aborted)), "gnm_p2p_status")
        assert aborted.value == 0, "a peer exchange gave up waiting"

        class _C(object):                      # what ops.bn_finalize / bn_bwd_coeffs expect as p2p
            def __init__(self, st):
                self.st = st

            def ref(self):
                return ctypes.byref(self.st)

        f, count = 64, 1000.0
        parts = [torch.rand(2 * f, dtype=torch.float64, device=DEV) * 500 + torch.cat(
            [torch.zeros(f, dtype=torch.float64, device=DEV), torch.full((f,), 400.0, dtype=torch.float64, device=DEV)])
            for _ in range(world)]
        gamma, beta = torch.rand(f, device=DEV) + 0.5, torch.randn(f, device=DEV)
        ref_out = [torch.empty(f, device=DEV) for _ in range(4)]
        ops.bn_finalize(parts[0] + parts[1], count, gamma, beta, 1e-5, 0.0, None, None, None, *ref_out)
        outs = [[torch.empty(f, device=DEV) for _ in range(4)] for _ in range(world)]
        work = [p.clone() for p in parts]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ops.bn_finalize(work[r], count, gamma, beta, 1e-5, 0.0, None, None, None, *outs[r], p2p=_C(comms[r]))
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(work[r], parts[0] + parts[1]), "bn_finalize leaves the global sums in place"
            for a, b in zip(outs[r], ref_out):
                assert torch.equal(a, b), "fused bn_finalize == bn_finalize on pre-summed statistics"
        mean, rstd = torch.randn(f, device=DEV), torch.rand(f, device=DEV) + 0.5
        ref_coef = torch.empty(3, f, device=DEV)
        ops.bn_bwd_coeffs(parts[0] + parts[1], count, gamma, mean, rstd, ref_coef)
        coefs = [torch.empty(3, f, device=DEV) for _ in range(world)]
        work = [p.clone() for p in parts]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ops.bn_bwd_coeffs(work[r], count, gamma, mean, rstd, coefs[r], p2p=_C(comms[r]))
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(coefs[r], ref_coef), "fused bn_bwd_coeffs"
        glib.check(L.gnm_p2p_status(ctypes.byref(aborted)), "gnm_p2p_status")
        assert aborted.value == 0
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            L.gnm_p2p_close(b, 1)


@pytest.mark.parametrize("n,b,f,mode,use_eps", [(400, 333, 64, 0, False), (400, 20, 64, 0, True), (48, 7, 32, 1, True),
                                                (12, 3, 8, 0, False), (416, 150, 64, 1, False)])
def test_aggregate_dense_table_shared_rows_and_column_stats(n, b, f, mode, use_eps):
    """Layer 0 with one table shared by all graphs (gnm_aggregate_dense_table): the B planes are converted once per CTA
    (b > 148 makes the persistent CTAs reuse them over several graphs); result and fused BatchNorm statistics against
    the general kernels (gnm_aggregate_dense with a per-row map + gnm_col_stats) and the fp64 stand-in."""
    rng = np.random.default_rng(n + b)
    counts = [n] * b
    ems = [rand_graph_edges(rng, n, 0.3) for _ in counts]
    if mode != 0:
        ring = np.stack([np.arange(n), (np.arange(n) + 1) % n])
        ems = [np.unique(np.concatenate([e, ring, ring[::-1]], 1), axis=1) for e in ems]
        ems = [e[:, e[0] != e[1]] for e in ems]
    e, eo, no = build_inputs(ems, counts)
    m = n * b
    self_loops = not use_eps
    rp, ci, _ = ops.csr_build(e, eo, no, b, n, m, self_loops, False)
    rpl, cil, _ = ops.csr_build(e, eo, no, b, n, m, self_loops, True)
    bm, dup, addr, _ = _bitmaps(rpl, cil, no, counts)
    torch.manual_seed(f + b)
    table = torch.randn(n + 5, f, device=DEV) * 3
    tags1 = torch.randperm(n + 5)[:n].to(torch.int32).to(DEV)            # one graph's injective tag sequence
    tags = tags1.repeat(b)
    eps = torch.tensor([0.3], device=DEV) if use_eps else None
    bias = torch.randn(f, device=DEV)
    got = torch.full((m, f), float("nan"), device=DEV)
    st = torch.zeros(2 * f, dtype=torch.float64, device=DEV)
    assert ops.aggregate_dense_table(addr, no, rp, b, n, table, tags1, got, mode, eps, bias, st)
    assert not ops.aggregate_tc_status(), "tcgen05 kernel hit a barrier timeout"
    want = torch.empty(m, f, device=DEV)
    ops.aggregate_dense(addr, no, rp, b, n, table, tags, want, mode, eps, bias, impl=2)
    assert_close(got, want, 1e-6, "shared table vs per-row map")
    st2 = torch.zeros(2 * f, dtype=torch.float64, device=DEV)
    ops.col_stats(want, st2)
    assert_close(st[:f], st2[:f], 1e-5, "fused column sums")
    assert_close(st[f:], st2[f:], 1e-5, "fused column sums of squares")
    ref = torch.empty(m, f)
    emul_ops.aggregate(rp.cpu(), ci.cpu(), table.cpu(), tags.cpu(), ref, mode, eps.cpu() if use_eps else None, bias.cpu())
    assert_close(got, ref, TOL, "shared table vs fp64")
    assert_close(st[:f], ref.double().sum(0), 1e-5, "column sums vs fp64")
    assert_close(st[f:], (ref.double() ** 2).sum(0), 1e-5, "column sums of squares vs fp64")
    # without statistics
    got2 = torch.empty(m, f, device=DEV)
    assert ops.aggregate_dense_table(addr, no, rp, b, n, table, tags1, got2, mode, eps, bias, None)
    assert torch.equal(got2, got)


@pytest.mark.parametrize("m,n_in,n_out,pro", [(8192, 64, 64, True), (40000, 64, 64, False), (5000, 32, 48, True)])
def test_linear_with_batchnorm_tail_equals_linear_then_bn_finalize(m, n_in, n_out, pro):
    """gnm_linear's BatchNorm tail (last CTA of the tcgen05 kernel runs gnm_bn_finalize's arithmetic on the statistics it
    just produced) against the two separate kernels: affine, saved statistics, running buffers, num_batches_tracked."""
    torch.manual_seed(m)
    x = torch.randn(m, n_in, device=DEV) * 2 + 0.3
    w = torch.randn(n_out, n_in, device=DEV) * 0.2
    b = torch.randn(n_out, device=DEV)
    sc = (torch.rand(n_in, device=DEV) + 0.5) if pro else None
    sh = torch.randn(n_in, device=DEV) if pro else None
    gamma, beta = torch.rand(n_out, device=DEV) + 0.5, torch.randn(n_out, device=DEV)

    def fresh():
        return (torch.zeros(2 * n_out, dtype=torch.float64, device=DEV), torch.empty(m, n_out, device=DEV),
                torch.empty(4, n_out, device=DEV), torch.full((n_out,), 0.25, device=DEV), torch.full((n_out,), 1.5, device=DEV),
                torch.tensor(3, dtype=torch.int64, device=DEV))
    st1, y1, o1, rm1, rv1, nbt1 = fresh()
    ops.linear(x, w, False, b, sc, sh, y1, st1)
    ops.bn_finalize(st1, float(m), gamma, beta, 1e-5, 0.1, rm1, rv1, nbt1, o1[0], o1[1], o1[2], o1[3])
    st2, y2, o2, rm2, rv2, nbt2 = fresh()
    tail = ops.BnTail(ops.BnTail.FINALIZE, float(m), gamma, o2[2], o2[3], beta=beta, eps=1e-5, momentum=0.1, running_mean=rm2,
                      running_var=rv2, nbt=nbt2, scale=o2[0], shift=o2[1])
    before = ops.launch_counts()
    assert ops.linear(x, w, False, b, sc, sh, y2, st2, tail) is True
    ran = {k: v - before[k] for k, v in ops.launch_counts().items()}
    assert ran["linear_tc"] == 1 and ran["other"] == 0, ran            # one kernel, no separate finalize
    assert torch.equal(y1, y2)
    assert_close(o2, o1, 1e-6, "scale / shift / mean / rstd")
    assert_close(rm2, rm1, 1e-6, "running_mean")
    assert_close(rv2, rv1, 1e-6, "running_var")
    assert int(nbt2) == int(nbt1) == 4
    # the ticket counter is left at zero: a second launch behaves the same
    st3, y3, o3, rm3, rv3, nbt3 = fresh()
    tail3 = ops.BnTail(ops.BnTail.FINALIZE, float(m), gamma, o3[2], o3[3], beta=beta, eps=1e-5, momentum=0.1, running_mean=rm3,
                       running_var=rv3, nbt=nbt3, scale=o3[0], shift=o3[1])
    assert ops.linear(x, w, False, b, sc, sh, y3, st3, tail3) is True
    assert_close(o3, o1, 1e-6, "second launch")
    # small problems run on the FFMA kernel, which has no tail: the call reports it and the Linear is still computed
    xs = x[:100].contiguous()
    st4, y4 = torch.zeros(2 * n_out, dtype=torch.float64, device=DEV), torch.empty(100, n_out, device=DEV)
    assert ops.linear(xs, w, False, b, sc, sh, y4, st4, tail3) is False
    assert_close(y4, y1[:100], 2e-5, "fallback result")
