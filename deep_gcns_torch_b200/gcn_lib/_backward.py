"""Backward glue of the autograd nodes of the drop-in modules: unpack what forward saved,
call the native gradient kernels (dgcn_graph_conv_backward / dgcn_genconv_aggregate_backward),
hand the gradients back in the order of the forward's inputs."""
from .. import _native


def graph_conv_backward(ctx, grad_out):
    """Inputs of _GraphConvFn.forward: (owner, x, edge_index, fused, weight, bias, prelu, bn_w, bn_b)."""
    (x,) = ctx.saved_tensors
    owner, prm = ctx.owner, ctx.prm
    need = ctx.needs_input_grad
    g = _native.graph_conv_backward(owner._conv, x, prm, grad_out, edge_index=ctx.edge_index,
                                    nbr=None if ctx.edge_index is not None else ctx.nbr, need_x=need[1])
    conv = owner.nn[0]
    gx = g["x"].view(x.shape[0], x.shape[1], x.shape[2], 1) if need[1] else None
    gw = g["weight"].view_as(conv.weight) if need[4] else None
    gb = g["bias"] if (need[5] and g["bias"] is not None) else None
    gp = g["prelu"] if (need[6] and g["prelu"] is not None) else None
    gbw = g["bn_weight"] if (need[7] and g["bn_weight"] is not None) else None
    gbb = g["bn_bias"] if (need[8] and g["bn_bias"] is not None) else None
    return None, gx, None, None, gw, gb, gp, gbw, gbb


def genconv_aggregate_backward(ctx, grad_out):
    """Inputs of _AggregateFn.forward: (owner, csr, raw, residual, x, edge_attr, t, p, y, msg_scale)."""
    import torch
    x, edge_attr = ctx.saved_tensors
    owner = ctx.owner
    t, p, y, msg_scale = ctx.scalars
    prm, keep = _native.genconv_params(ctx.aggr, t, p, y, getattr(owner, "eps", 1e-7), msg_scale,
                                       add_residual=ctx.residual)
    prm.raw_message = int(ctx.raw)
    need = ctx.needs_input_grad
    gsrc, gdst, gea, gsc = _native.genconv_aggregate_backward(
        x, None if ctx.raw else x, ctx.csr, prm, grad_out, edge_attr,
        softmax_grad=getattr(owner, "learn_t", False), need_edge_attr=need[5])
    gx = None
    if need[4]:
        gx = gsrc if gdst is None else gsrc + gdst
    def scalar_grad(i, v, idx):
        return gsc[idx:idx + 1].clone() if (need[i] and torch.is_tensor(v)) else None
    return (None, None, None, None, gx, gea if need[5] else None, scalar_grad(6, t, 0), scalar_grad(7, p, 1),
            scalar_grad(8, y, 2), scalar_grad(9, msg_scale, 3))
