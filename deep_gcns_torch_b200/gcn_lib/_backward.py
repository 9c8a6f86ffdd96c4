"""Backward glue for the autograd nodes of the drop-in modules (filled in by the
D5 / S6 milestone: dgcn_graph_conv_backward, dgcn_genconv_aggregate_backward)."""


def graph_conv_backward(ctx, grad_out):
    raise NotImplementedError("dense graph-conv backward is not available in this build")


def genconv_aggregate_backward(ctx, grad_out):
    raise NotImplementedError("GENConv aggregate backward is not available in this build")
