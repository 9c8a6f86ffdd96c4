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


def act_layer(act, inplace=False, neg_slope=0.2, n_prelu=1):
    """torch_nn.py:9-21."""
    kind = act.lower()
    if kind == "relu":
        return nn.ReLU(inplace)
    if kind == "leakyrelu":
        return nn.LeakyReLU(neg_slope, inplace)
    if kind == "prelu":
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    raise NotImplementedError("activation layer [%s] is not found" % kind)


def norm_layer(norm, nc):
    """torch_nn.py:24-33."""
    kind = norm.lower()
    if kind == "batch":
        return nn.BatchNorm2d(nc, affine=True)
    if kind == "instance":
        return nn.InstanceNorm2d(nc, affine=False)
    raise NotImplementedError("normalization layer [%s] is not found" % kind)


def _wanted(name):
    return name is not None and name.lower() != "none"


class MLP(nn.Sequential):
    """torch_nn.py:36-45 (Linear -> act -> norm); not used by the hot path."""

    def __init__(self, channels, act="relu", norm=None, bias=True):
        layers = []
        for c_in, c_out in zip(channels[:-1], channels[1:]):
            layers.append(nn.Linear(c_in, c_out, bias))
            if _wanted(act):
                layers.append(act_layer(act))
            if _wanted(norm):
                layers.append(norm_layer(norm, channels[-1]))
        super().__init__(*layers)


class BasicConv(nn.Sequential):
    """torch_nn.py:48-72."""

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0.):
        layers = []
        for c_in, c_out in zip(channels[:-1], channels[1:]):
            layers.append(nn.Conv2d(c_in, c_out, 1, bias=bias))
            if _wanted(act):
                layers.append(act_layer(act))
            if _wanted(norm):
                layers.append(norm_layer(norm, channels[-1]))
            if drop > 0:
                layers.append(nn.Dropout2d(drop))
        super().__init__(*layers)
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.InstanceNorm2d):
                # the reference dereferences the (absent) affine weight here and raises
                # AttributeError (torch_nn.py:71): 'instance' cannot be built through BasicConv.
                m.weight.data.fill_(1)
                m.bias.data.zero_()


def batched_index_select(x, idx):
    """torch_nn.py:75-96: x (B,C,N,1), idx (B,N,k) -> (B,C,N,k).  Kept for callers
    outside the hot path; the graph convolutions never materialise this tensor."""
    B, C, N = x.shape[:3]
    k = idx.shape[-1]
    flat = (idx + torch.arange(B, device=idx.device).view(B, 1, 1) * N).reshape(-1)
    rows = x.squeeze(-1).transpose(1, 2).reshape(B * N, C).index_select(0, flat)
    return rows.view(B, N, k, C).permute(0, 3, 1, 2).contiguous()
