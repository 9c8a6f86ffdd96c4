"""Dense nn helpers - API of the reference's gcn_lib/dense/torch_nn.py.

`BasicConv` stays a plain torch Sequential (1x1 Conv2d -> act -> norm -> Dropout2d,
torch_nn.py:48-72): standing alone (fusion / prediction heads) it is outside the
hot path; inside EdgeConv2d / MRConv2d it is the parameter container whose tensors
the CUDA kernels read, so its state_dict keys (`nn.0.weight`, `nn.2.running_mean`,
...) are exactly the reference's.
"""
import torch
from torch import nn

__all__ = ["act_layer", "norm_layer", "MLP", "BasicConv", "batched_index_select"]

# name -> factory(inplace, negative slope, PReLU parameter count)          (torch_nn.py:9-21)
_ACTIVATIONS = {
    "relu": lambda inplace, slope, n: nn.ReLU(inplace),
    "leakyrelu": lambda inplace, slope, n: nn.LeakyReLU(slope, inplace),
    "prelu": lambda inplace, slope, n: nn.PReLU(num_parameters=n, init=slope),
}
# name -> factory(channels)                                                 (torch_nn.py:24-33)
_NORMS = {
    "batch": lambda nc: nn.BatchNorm2d(nc, affine=True),
    "instance": lambda nc: nn.InstanceNorm2d(nc, affine=False),
}


def _lookup(table, name, what):
    try:
        return table[name.lower()]
    except KeyError:
        raise NotImplementedError("%s layer [%s] is not found" % (what, name.lower())) from None


def act_layer(act, inplace=False, neg_slope=0.2, n_prelu=1):
    return _lookup(_ACTIVATIONS, act, "activation")(inplace, neg_slope, n_prelu)


def norm_layer(norm, nc):
    return _lookup(_NORMS, norm, "normalization")(nc)


def _named(option):
    return option is not None and option.lower() != "none"


def _stages(channels, unit, act, norm, drop=0.):
    """unit(c_in, c_out) -> act -> norm(channels[-1]) [-> Dropout2d] for consecutive channel pairs, in
    the reference's module order (it fixes the state_dict indices and the order of RNG draws)."""
    for c_in, c_out in zip(channels[:-1], channels[1:]):
        yield unit(c_in, c_out)
        if _named(act):
            yield act_layer(act)
        if _named(norm):
            yield norm_layer(norm, channels[-1])
        if drop > 0:
            yield nn.Dropout2d(drop)


class MLP(nn.Sequential):
    """torch_nn.py:36-45; not used by the hot path."""

    def __init__(self, channels, act="relu", norm=None, bias=True):
        super().__init__(*_stages(channels, lambda i, o: nn.Linear(i, o, bias), act, norm))


class BasicConv(nn.Sequential):
    """torch_nn.py:48-72."""

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0.):
        super().__init__(*_stages(channels, lambda i, o: nn.Conv2d(i, o, 1, bias=bias), act, norm, drop))
        self.reset_parameters()

    def reset_parameters(self):
        """Kaiming-normal 1x1 weights, zero biases, unit / zero affine norms (torch_nn.py:64-72)."""
        for layer in self.modules():
            if isinstance(layer, nn.Conv2d):
                nn.init.kaiming_normal_(layer.weight)
                if layer.bias is not None:
                    nn.init.zeros_(layer.bias)
            elif isinstance(layer, (nn.BatchNorm2d, nn.InstanceNorm2d)):
                # InstanceNorm2d(affine=False) has no weight: like the reference (torch_nn.py:71) this
                # raises AttributeError, i.e. 'instance' cannot be built through BasicConv.
                layer.weight.data.fill_(1)
                layer.bias.data.zero_()


def batched_index_select(x, idx):
    """torch_nn.py:75-96: x (B,C,N,1), idx (B,N,k) -> (B,C,N,k).  Kept for callers
    outside the hot path; the graph convolutions never materialise this tensor."""
    B, C, N = x.shape[:3]
    k = idx.shape[-1]
    flat = (idx + torch.arange(B, device=idx.device).view(B, 1, 1) * N).reshape(-1)
    rows = x.squeeze(-1).transpose(1, 2).reshape(B * N, C).index_select(0, flat)
    return rows.view(B, N, k, C).permute(0, 3, 1, 2).contiguous()
