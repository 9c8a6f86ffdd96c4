"""Gradients of the dense graph convolutions (dgcn_graph_conv_backward) against torch
autograd through the oracle restatement of the reference forward."""
import pytest
import torch

from oracle import dense as od

pytestmark = pytest.mark.gpu

CASES = [
    # B, C, Co, N, k, d, conv, act, norm, bias, train
    (2, 8, 12, 96, 5, 1, "edge", "relu", "batch", True, False),
    (2, 8, 12, 96, 5, 2, "edge", "leakyrelu", "batch", True, True),
    (1, 6, 40, 130, 4, 1, "edge", "prelu", None, False, False),
    (2, 8, 12, 96, 5, 1, "mr", "relu", "batch", True, False),
    (2, 16, 24, 128, 6, 2, "mr", "leakyrelu", "batch", True, True),
    (1, 6, 10, 70, 3, 1, "mr", "prelu", None, True, False),
    (2, 64, 64, 256, 20, 1, "edge", "relu", "batch", True, True),
    (2, 64, 64, 256, 20, 1, "mr", "relu", "batch", True, True),
]


@pytest.mark.parametrize("cfg", CASES)
def test_grads_match_autograd_of_oracle(cfg):
    from deep_gcns_torch_b200.gcn_lib import dense as D
    B, C, Co, N, k, d, conv, act, norm, bias, train = cfg
    g = torch.Generator().manual_seed(sum(cfg[:6]))
    torch.manual_seed(0)
    mod = D.DynConv2d(C, Co, k, d, conv, act, norm, bias)
    for m in mod.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = torch.randn(Co, generator=g) * 0.5 + 0.8
            m.weight.data[::3] *= -1                     # negative gammas: the max turns into a min
            m.bias.data = torch.randn(Co, generator=g) * 0.2
            m.running_mean.data = torch.randn(Co, generator=g) * 0.3
            m.running_var.data = torch.rand(Co, generator=g) + 0.4
    x = torch.randn(B, C, N, 1, generator=g)
    wgt = torch.randn(B, Co, N, 1, generator=g)
    ei = od.dilated_knn_graph(x, k, d)

    # oracle: torch autograd through the reference formulation (fp64 for a clean gradient reference)
    p = od.params_from_module(mod.gconv.nn, dtype=torch.float64)
    leaves = {"x": x.double().requires_grad_(True), "weight": p["weight"].requires_grad_(True)}
    if "bias" in p:
        leaves["bias"] = p["bias"].requires_grad_(True)
    if "slope" in p:
        leaves["slope"] = p["slope"].requires_grad_(True)
    if "norm" in p:
        leaves["bn_w"] = p["norm"]["weight"].requires_grad_(True)
        leaves["bn_b"] = p["norm"]["bias"].requires_grad_(True)
    y_ref = od.graph_conv(leaves["x"], ei, p, conv, act, norm, train)
    (y_ref * wgt.double()).sum().backward()

    mod = mod.cuda().train(train)
    xc = x.cuda().requires_grad_(True)
    for static in (True, False):
        mod.zero_grad()
        xc.grad = None
        y = mod(xc, ei.cuda()) if static else mod(xc)
        torch.testing.assert_close(y.detach().cpu(), y_ref.detach().float(), rtol=1e-3, atol=1e-4)
        (y * wgt.cuda()).sum().backward()
        nn = mod.gconv.nn
        got = {"x": xc.grad, "weight": nn[0].weight.grad, "bias": nn[0].bias.grad if bias else None}
        for m in nn:
            if isinstance(m, torch.nn.PReLU):
                got["slope"] = m.weight.grad
            if isinstance(m, torch.nn.BatchNorm2d):
                got["bn_w"], got["bn_b"] = m.weight.grad, m.bias.grad
        for name, leaf in leaves.items():
            ref = leaf.grad.float().reshape(got[name].shape)
            scale = ref.abs().max().clamp_min(1e-6)
            err = (got[name].cpu() - ref).abs().max() / scale
            assert err < 2e-3, (name, static, float(err))
