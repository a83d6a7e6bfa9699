"""DGI `Discriminator` with the reference's constructor, forward signature and state_dict
(reference: models/discriminator.py:5-38), scored by libgnm.

`nn.Bilinear(n_h, n_h, 1)` evaluates h^T W c + b. Because every row of one graph is scored
against the same summary c_g, that is <h, u_g> + b with u_g = W c_g: B small mat-vecs plus
one row-dot per node, instead of ATen's `_trilinear` expansion over all M rows (73 % of the
reference's CPU step).
"""
import torch
import torch.nn as nn

from .. import engine as _engine
from .. import ops as _ops


class RowDotScore(torch.autograd.Function):
    """out[r] = <h[r], u[r // rows_per_graph]> + bias  (discriminator.py:28-29)."""

    @staticmethod
    def forward(ctx, h, u, bias, rows_per_graph):
        h_c, u_c = h.detach().contiguous(), u.detach().contiguous()
        out = torch.empty(h_c.shape[0], dtype=torch.float32, device=h.device)
        _ops.rowdot_score(h_c, u_c, rows_per_graph, bias.detach(), None, out)
        ctx.save_for_backward(h_c, u_c)
        ctx.rows_per_graph = rows_per_graph
        return out.unsqueeze(1)

    @staticmethod
    def backward(ctx, dout):
        h, u = ctx.saved_tensors
        n = ctx.rows_per_graph
        b = u.shape[0]
        d = dout.reshape(-1)[:b * n].reshape(b, n, 1)
        dh = du = None
        if ctx.needs_input_grad[0]:
            dh = torch.zeros_like(h)
            dh[:b * n] = (d * u.unsqueeze(1)).reshape(b * n, -1)
        if ctx.needs_input_grad[1]:
            du = torch.bmm(d.transpose(1, 2), h[:b * n].reshape(b, n, -1)).squeeze(1)
        return dh, du, dout.sum().reshape(1), None


class Discriminator(nn.Module):
    def __init__(self, n_h):
        super().__init__()
        self.f_k = nn.Bilinear(n_h, n_h, 1)
        self.weights_init(self.f_k)

    def weights_init(self, m):
        if isinstance(m, nn.Bilinear):
            nn.init.xavier_uniform_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()

    def summary_vectors(self, c):
        """u_g = W c_g for every graph summary row of c: [B, n_h]."""
        return c @ self.f_k.weight[0].t()

    def forward(self, c, h_pl, h_mi, s_bias1=None, s_bias2=None):
        _engine.require_cuda(h_pl.device)
        rows = h_pl.shape[0] // c.shape[0]        # discriminator.py:23-26: each c_g covers M // B rows
        if rows * c.shape[0] != h_pl.shape[0] or h_mi.shape[0] != h_pl.shape[0]:
            raise RuntimeError("Discriminator: %d rows cannot be split evenly over %d summaries"
                               % (h_pl.shape[0], c.shape[0]))
        u = self.summary_vectors(c)
        sc_1 = RowDotScore.apply(h_pl, u, self.f_k.bias, rows)
        sc_2 = RowDotScore.apply(h_mi, u, self.f_k.bias, rows)
        if s_bias1 is not None:
            sc_1 = sc_1 + s_bias1
        if s_bias2 is not None:
            sc_2 = sc_2 + s_bias2
        return torch.cat((sc_1, sc_2), 0)
