"""GPU parity of the step-remainder kernels (gnm_train.cu) against the torch operators the reference's
driver calls: nn.Linear heads + F.dropout + nn.CrossEntropyLoss (graphcnn.py:228-231, main.py:16,35),
nn.BCEWithLogitsLoss (main.py:17,34), the sigmoid / bilinear glue of the Discriminator and torch.optim.Adam
(main.py:136). fp32 tolerances, scaled by each tensor's max-abs."""
import numpy as np
import pytest
import torch

from helpers import assert_close
from graph_neural_mapping_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("b,layers,feat,classes,drop", [(1024, 5, 64, 2, 0.5), (7, 3, 8, 2, 0.0), (100, 2, 128, 3, 0.3),
                                                        (1, 1, 12, 2, 0.0), (8192, 5, 64, 2, 0.5)])
def test_heads_ce_matches_torch_autograd(b, layers, feat, classes, drop):
    torch.manual_seed(b + layers)
    g_f = (torch.randn(b, layers * feat, device=DEV) * 3.0).requires_grad_(True)
    ws = [torch.randn(classes, feat, device=DEV).mul_(0.1).requires_grad_(True) for _ in range(layers)]
    bs = [torch.randn(classes, device=DEV).requires_grad_(True) for _ in range(layers)]
    labels = torch.randint(0, classes, (b,), device=DEV)
    mask = None
    if drop > 0:
        mask = torch.empty(layers, b, classes, device=DEV).bernoulli_(1 - drop).mul_(1 / (1 - drop))
    score = 0
    for l in range(layers):
        y = torch.nn.functional.linear(g_f[:, l * feat:(l + 1) * feat], ws[l], bs[l])
        score = score + (y * mask[l] if mask is not None else y)
    loss = torch.nn.functional.cross_entropy(score, labels)
    loss.backward()
    c_logit = torch.empty(b, classes, device=DEV)
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    d_gf = torch.empty(b, layers * feat, device=DEV)
    dws = [torch.full_like(w, 7.0) for w in ws]          # outputs are written, not accumulated
    dbs = [torch.full_like(x, 7.0) for x in bs]
    wsb = torch.empty(ops.heads_ce_workspace(b, layers, feat, classes), device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    for rep in range(2):                                  # the counter is left at zero: a second call works unchanged
        acc.zero_()
        ops.heads_ce(g_f.detach(), [w.detach() for w in ws], [x.detach() for x in bs], mask, labels, 1.0 / b, c_logit, acc,
                     d_gf, dws, dbs, wsb, counter)
        assert int(counter) == 0
        assert_close(c_logit, score.detach(), 2e-5, "c_logit")
        assert_close(acc[0], loss.detach(), 2e-5, "CE loss")
        assert_close(d_gf, g_f.grad, 2e-5, "d g_f")
        for l in range(layers):
            assert_close(dws[l], ws[l].grad, 5e-5, "dW %d" % l)
            assert_close(dbs[l], bs[l].grad, 5e-5, "db %d" % l)
    first = [d.clone() for d in dws]
    ops.heads_ce(g_f.detach(), [w.detach() for w in ws], [x.detach() for x in bs], mask, labels, 1.0 / b, c_logit, acc, d_gf,
                 dws, dbs, wsb, counter)
    assert all(torch.equal(a, c) for a, c in zip(first, dws)), "head gradients are deterministic (fixed-order reduction)"


@pytest.mark.parametrize("m", [1, 37, 409600])
def test_bce_logits_matches_torch(m):
    torch.manual_seed(m)
    x = (torch.randn(2 * m, 1, device=DEV) * 6.0).requires_grad_(True)
    if m > 1:
        x.data[0] = 60.0
        x.data[-1] = -60.0                               # the stable form must survive saturated scores
    y = torch.cat([torch.ones(m, 1), torch.zeros(m, 1)], 0).to(DEV)          # main.py:32
    beta = 0.05
    loss = beta * torch.nn.functional.binary_cross_entropy_with_logits(x, y)
    loss.backward()
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    dx = torch.empty(2 * m, device=DEV)
    w = beta / (2.0 * m)
    ops.bce_logits(x.detach().view(-1), m, w, w, acc, dx)
    assert_close(acc[0], loss.detach(), 1e-5, "BCE loss")
    assert_close(dx, x.grad.view(-1), 1e-5, "d logits")
    acc.zero_()
    ops.bce_logits(x.detach().view(-1), m, w, w, acc, None)                   # loss only
    assert_close(acc[0], loss.detach(), 1e-5, "BCE loss, no gradient")


@pytest.mark.parametrize("b,lf", [(1024, 320), (5, 24), (130, 100)])
def test_discriminator_glue_matches_torch(b, lf):
    torch.manual_seed(b)
    g_f = torch.randn(b, lf, device=DEV) * 2.0
    w = torch.randn(lf, lf, device=DEV) * 0.1
    c = torch.empty(b, lf, device=DEV)
    u = torch.empty(b, lf, device=DEV)
    ops.small_gemm(g_f, (lf, 1), w, (1, lf), u, b, lf, lf, sigmoid_a_out=c)
    c_ref = torch.sigmoid(g_f.double())
    assert_close(c, c_ref, 1e-6, "c = sigmoid(g_f)")
    assert_close(u, c_ref @ w.double().t(), 2e-5, "u = c W^T")
    du = torch.randn(b, lf, device=DEV)
    dw = torch.empty(lf, lf, device=DEV)
    ops.small_gemm(du, (1, lf), c, (lf, 1), dw, lf, lf, b)
    assert_close(dw, du.double().t() @ c.double(), 2e-5, "dW = du^T c")
    dg_heads = torch.randn(b, lf, device=DEV)
    dg = torch.empty(b, lf, device=DEV)
    ops.small_gemm(du, (lf, 1), w, (lf, 1), dg, b, lf, lf, dsig_s=c, dsig_add=dg_heads)
    ref = dg_heads.double() + (du.double() @ w.double()) * c.double() * (1 - c.double())
    assert_close(dg, ref, 2e-5, "d g_f")
    ops.small_gemm(du, (lf, 1), w, (lf, 1), dg, b, lf, lf, dsig_s=c)
    assert_close(dg, ref - dg_heads.double(), 2e-5, "d g_f without the heads term")
    # gradient reaching the shuffled rows: permutation (single process) and a slice of one (data parallel shard)
    s2 = torch.randn(b, device=DEV)
    for n_neg, idx in [(b, torch.randperm(b)), (3 * b, torch.randperm(3 * b)[:b]), (b, torch.zeros(b, dtype=torch.int64))]:
        neg = idx.to(DEV).to(torch.int32)
        if int(neg.max()) == 0 and b > 64:
            continue                                      # more than 64 graphs naming one row is outside the contract
        d_neg = torch.full((n_neg, lf), 9.0, device=DEV)
        ops.dgi_neg_grad(neg, s2, u, d_neg)
        want = torch.zeros(n_neg, lf, dtype=torch.float64, device=DEV)
        want.index_add_(0, neg.long(), (s2.unsqueeze(1) * u).double())
        assert_close(d_neg, want, 2e-5, "d_neg")


def test_adam_step_matches_torch_adam_over_several_steps_and_lr_changes():
    torch.manual_seed(0)
    shapes = [(5,), (1, 320, 320), (1,), (64, 400), (64,), (64, 64), (2, 64), (2,), (3, 7)]
    ref_p = [torch.randn(*s, device=DEV).requires_grad_(True) for s in shapes]
    own_p = [p.detach().clone() for p in ref_p]
    opt = torch.optim.Adam(ref_p, lr=0.005)
    offs, total = [], 0
    for p in own_p:
        offs.append(total)
        total += (p.numel() + 3) // 4 * 4
    m = torch.zeros(total, device=DEV)
    v = torch.zeros(total, device=DEV)
    step = torch.zeros(2, device=DEV)
    lr = torch.tensor(0.005, device=DEV)
    terms = torch.tensor([1.5, 0.25], dtype=torch.float64, device=DEV)
    loss_out = torch.zeros(1, device=DEV)
    for it in range(12):
        grads = [torch.randn_like(p) * (10.0 ** (it % 4 - 2)) for p in ref_p]
        if it == 6:
            for grp in opt.param_groups:
                grp["lr"] = 0.001
            lr.fill_(0.001)
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        opt.step()
        ops.adam_step(own_p, grads, offs, m, v, step, lr, 0.9, 0.999, 1e-8, 0.0, 1.0, loss_terms=terms, loss_out=loss_out)
        assert float(step[0]) == it + 1
        assert abs(float(loss_out) - 1.75) < 1e-6
        for i, (a, r) in enumerate(zip(own_p, ref_p)):
            assert_close(a, r.detach(), 2e-6, "param %d after step %d" % (i, it))
    for i, p in enumerate(ref_p):
        st = opt.state[p]
        n = p.numel()
        assert_close(m[offs[i]:offs[i] + n].view_as(p), st["exp_avg"], 1e-5, "exp_avg %d" % i)
        assert_close(v[offs[i]:offs[i] + n].view_as(p), st["exp_avg_sq"], 1e-5, "exp_avg_sq %d" % i)
    # grad_scale: gradients summed over `world` ranks are averaged on the fly
    a = [own_p[3].clone()]
    b2 = [own_p[3].clone()]
    g = torch.randn_like(a[0])
    ma, va, sa = torch.zeros(a[0].numel(), device=DEV), torch.zeros(a[0].numel(), device=DEV), torch.zeros(2, device=DEV)
    mb, vb, sb = ma.clone(), va.clone(), sa.clone()
    ops.adam_step(a, [g * 4.0], [0], ma, va, sa, lr, 0.9, 0.999, 1e-8, 0.0, 0.25)
    ops.adam_step(b2, [g], [0], mb, vb, sb, lr, 0.9, 0.999, 1e-8, 0.0, 1.0)
    assert_close(a[0], b2[0], 1e-6, "grad_scale")


@pytest.mark.parametrize("b,layers,feat,classes,drop", [(1024, 5, 64, 2, 0.5), (3, 2, 8, 2, 0.0), (77, 3, 128, 4, 0.25)])
def test_heads_forward_backward_halves_match_torch(b, layers, feat, classes, drop):
    """gnm_heads_fwd / gnm_heads_bwd (the model's _HeadsFunction: the caller's own loss sits in between) against
    nn.Linear + masks + autograd, and the module-level function against the reference's per-layer loop."""
    from graph_neural_mapping_b200.models.graphcnn import _HeadsFunction
    torch.manual_seed(b)
    g_f = (torch.randn(b, layers * feat, device=DEV) * 2.0).requires_grad_(True)
    ws = [torch.randn(classes, feat, device=DEV).mul_(0.2).requires_grad_(True) for _ in range(layers)]
    bs = [torch.randn(classes, device=DEV).requires_grad_(True) for _ in range(layers)]
    mask = torch.empty(layers, b, classes, device=DEV).bernoulli_(1 - drop).mul_(1 / (1 - drop)) if drop > 0 else None
    score = 0
    for l in range(layers):
        y = torch.nn.functional.linear(g_f[:, l * feat:(l + 1) * feat], ws[l], bs[l])
        score = score + (y * mask[l] if mask is not None else y)
    go = torch.randn(b, classes, device=DEV)
    ref = torch.autograd.grad(score, [g_f] + ws + bs, go)
    g2 = g_f.detach().clone().requires_grad_(True)
    ws2 = [w.detach().clone().requires_grad_(True) for w in ws]
    bs2 = [x.detach().clone().requires_grad_(True) for x in bs]
    out = _HeadsFunction.apply(g2, mask, layers, *(ws2 + bs2))
    assert_close(out, score.detach(), 2e-5, "c_logit")
    got = torch.autograd.grad(out, [g2] + ws2 + bs2, go)
    for a, r, nm in zip(got, ref, ["d g_f"] + ["dW"] * layers + ["db"] * layers):
        assert_close(a, r, 5e-5, nm)
