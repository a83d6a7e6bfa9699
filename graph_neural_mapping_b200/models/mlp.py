"""`MLP` with the reference's constructor, attribute names and state_dict layout
(reference: models/mlp.py:6-49), computed by the libgnm kernels.

Inside `GIN_InfoMaxReg` the MLP's Linear/BatchNorm/ReLU chain is executed by the fused
engine (`engine.run_forward`), which reads this module's parameters directly. Calling the
module on its own (`mlp(x)`) runs the same kernels through `MLPFunction`.
"""
import torch
import torch.nn as nn

from .. import engine as _engine
from .. import ops as _ops


class MLPFunction(torch.autograd.Function):
    """x -> [Linear -> BatchNorm1d -> ReLU] x (k-1) -> Linear, forward and backward on libgnm."""

    @staticmethod
    def forward(ctx, mlp, x, *params):
        k = mlp.num_layers
        dev = x.device
        x = x.detach().contiguous()
        m = x.shape[0]
        training = mlp.training
        ws = [p.detach() for p in params[0::2][:k]]
        bs = [p.detach() for p in params[1::2][:k]]
        bnp = [p.detach() for p in params[2 * k:]]
        zs, affs, counts = [], [], []
        cur, sc, sh = x, None, None
        for j in range(k):
            z = torch.empty(m, ws[j].shape[0], dtype=torch.float32, device=dev)
            last = j == k - 1
            stats = None if last else torch.zeros(2 * ws[j].shape[0], dtype=torch.float64, device=dev)
            _ops.linear(cur, ws[j], False, bs[j], sc, sh, z, stats)
            zs.append(z)
            if not last:
                bn = mlp.batch_norms[j]
                buf = torch.empty(4, z.shape[1], dtype=torch.float32, device=dev)
                if training or not bn.track_running_stats:
                    track = training and bn.track_running_stats
                    _ops.bn_finalize(stats, float(m), bnp[2 * j], bnp[2 * j + 1], bn.eps, bn.momentum if track else 0.0,
                                     bn.running_mean if track else None, bn.running_var if track else None,
                                     bn.num_batches_tracked if track else None, buf[0], buf[1], buf[2], buf[3])
                    counts.append(float(m))
                else:
                    _ops.bn_eval_affine(bn.running_mean, bn.running_var, bnp[2 * j], bnp[2 * j + 1], bn.eps,
                                        buf[0], buf[1], buf[2], buf[3])
                    counts.append(0.0)
                affs.append(buf)
                cur, sc, sh = z, buf[0], buf[1]
        ctx.mlp, ctx.x, ctx.zs, ctx.affs, ctx.counts = mlp, x, zs, affs, counts
        ctx.ws, ctx.bnp = ws, bnp
        return zs[-1]

    @staticmethod
    def backward(ctx, dout):
        mlp, x, zs, affs, counts, ws, bnp = ctx.mlp, ctx.x, ctx.zs, ctx.affs, ctx.counts, ctx.ws, ctx.bnp
        k = mlp.num_layers
        dev = x.device
        m = x.shape[0]
        # relu_bn_bwd_reduce walks row segments ("graphs"); here: plain chunks of 256 rows
        seg = torch.cat([torch.arange(0, m, 256, dtype=torch.int32, device=dev),
                         torch.tensor([m], dtype=torch.int32, device=dev)])
        nseg = seg.numel() - 1
        dws, dbs, dgam, dbet = [None] * k, [None] * k, [None] * (k - 1), [None] * (k - 1)
        dz = dout.contiguous()
        for j in range(k - 1, -1, -1):
            dw = torch.zeros_like(ws[j])
            db = torch.zeros(ws[j].shape[0], dtype=torch.float32, device=dev)
            if j > 0:
                a = affs[j - 1]
                _ops.linear_wgrad(dz, zs[j - 1], a[0], a[1], dw, db)
                da = torch.empty(m, ws[j].shape[1], dtype=torch.float32, device=dev)
                _ops.linear(dz, ws[j], True, None, None, None, da, None)
                n = zs[j - 1].shape[1]
                dy = torch.empty(m, n, dtype=torch.float32, device=dev)
                stats = torch.zeros(2 * n, dtype=torch.float64, device=dev)
                _ops.relu_bn_bwd_reduce(zs[j - 1], a[0], a[1], a[2], a[3], da, None, None, None, None, None, 0, seg,
                                        nseg, dy, stats)
                dbet[j - 1] = stats[:n].to(torch.float32)
                dgam[j - 1] = stats[n:].to(torch.float32)
                use_batch = counts[j - 1] > 0
                _ops.bn_bwd_apply(zs[j - 1], a[2], a[3], bnp[2 * (j - 1)], stats if use_batch else None, counts[j - 1], dy)
                dz = dy
            else:
                _ops.linear_wgrad(dz, x, None, None, dw, db)
                dx = torch.empty(m, ws[0].shape[1], dtype=torch.float32, device=dev)
                _ops.linear(dz, ws[0], True, None, None, None, dx, None)
            dws[j], dbs[j] = dw, db
        out = []
        for j in range(k):
            out.extend([dws[j], dbs[j]])
        for j in range(k - 1):
            out.extend([dgam[j], dbet[j]])
        return (None, dx) + tuple(out)


class MLP(nn.Module):
    """Linear, or Linear -> (BatchNorm1d -> ReLU -> Linear) x (num_layers - 1)."""

    def __init__(self, num_layers, input_dim, hidden_dim, output_dim):
        super().__init__()
        if num_layers < 1:
            raise ValueError("number of layers should be positive!")
        self.num_layers = num_layers
        self.linear_or_not = num_layers == 1
        if self.linear_or_not:
            self.linear = nn.Linear(input_dim, output_dim)
        else:
            dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
            self.linears = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers))
            self.batch_norms = nn.ModuleList(nn.BatchNorm1d(hidden_dim) for _ in range(num_layers - 1))

    def _flat(self):
        if self.linear_or_not:
            return [self.linear.weight, self.linear.bias]
        ps = []
        for lin in self.linears:
            ps.extend([lin.weight, lin.bias])
        for bn in self.batch_norms:
            ps.extend([bn.weight, bn.bias])
        return ps

    def forward(self, x):
        _engine.require_cuda(x.device)
        return MLPFunction.apply(self, x, *self._flat())
