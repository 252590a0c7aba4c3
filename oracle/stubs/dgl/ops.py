"""dgl.ops.gspmm('mul','sum') leaf restated with torch CSR; see dgl/__init__.py."""
import torch

_cache = {}


def _csr_pair(g, vals):
    key = (g.src.data_ptr(), g.dst.data_ptr(), vals.data_ptr(), int(vals.numel()), g.num_nodes)
    hit = _cache.get(key)
    if hit is not None and hit[0] is vals:
        return hit[1], hit[2]
    n = g.num_nodes
    coo = torch.sparse_coo_tensor(torch.stack([g.dst, g.src]), vals.detach(), (n, n)).coalesce()
    a = coo.to_sparse_csr()
    at = coo.t().coalesce().to_sparse_csr()
    if len(_cache) > 8:
        _cache.clear()
    _cache[key] = (vals, a, at)
    return a, at


class _GSpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, at):
        ctx.at = at
        return torch.sparse.mm(a, x)

    @staticmethod
    def backward(ctx, gy):
        return torch.sparse.mm(ctx.at, gy.contiguous()), None, None


def gspmm(g, op, reduce_op, lhs_data=None, rhs_data=None):
    assert op == 'mul' and reduce_op == 'sum'
    a, at = _csr_pair(g, rhs_data)
    return _GSpMM.apply(lhs_data, a, at)
