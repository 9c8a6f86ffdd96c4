"""Gradients of the fused GENConv aggregate (dgcn_genconv_aggregate_backward) against torch
autograd through the oracle restatement of the reference forward (fp64)."""
import pytest
import torch

from oracle import sparse as osp

pytestmark = pytest.mark.gpu

CFGS = [
    dict(aggr="softmax_sg", t=0.1),
    dict(aggr="softmax", t=0.6, learn_t=True, msg_norm=True, learn_msg_scale=True),
    dict(aggr="softmax", t=1.0),
    dict(aggr="softmax_sum", t=0.4, learn_t=True, y=0.3, learn_y=True),
    dict(aggr="power", p=2.0, learn_p=True, msg_norm=True),
    dict(aggr="power_sum", p=1.5, learn_p=True, y=-0.2, learn_y=True),
    dict(aggr="add"),
    dict(aggr="mean", msg_norm=True),
    dict(aggr="max"),
]


@pytest.mark.parametrize("C", [24, 128, 200])
@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: c["aggr"] + ("_lt" if c.get("learn_t") else ""))
def test_grads_match_autograd_of_oracle(cfg, C):
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    g = torch.Generator().manual_seed(C)
    N, E = 260, 3000
    dst = torch.randint(0, N - 20, (E,), generator=g)
    dst[:700] = 5
    ei = torch.stack((torch.randint(0, N, (E,), generator=g), dst))
    x = torch.randn(N, C, generator=g)
    wgt = torch.randn(N, C, generator=g)
    torch.manual_seed(1)
    mod = S.GENConv(C, C, mlp_layers=1, norm="layer", **cfg)
    if mod.msg_norm is not None:
        mod.msg_norm.msg_scale.data.fill_(0.7)

    # fp64 oracle with autograd
    ref = __import__("copy").deepcopy(mod).double()
    xr = x.double().requires_grad_(True)
    scal = lambda name, d: getattr(ref, name, d)
    h = osp.genconv_pre_mlp(xr, ei, None, ref.aggr, scal("t", 1.0), scal("p", 1.0), scal("y", 0.0),
                            ref.msg_norm.msg_scale if ref.msg_norm is not None else None, ref.eps)
    if not getattr(ref, "learn_t", False) and ref.aggr in ("softmax", "softmax_sg", "softmax_sum"):
        # the reference computes the softmax weights under no_grad unless learn_t (torch_message.py:51-55)
        msg = osp.message(xr, ei, None, ref.eps)
        with torch.no_grad():
            z = msg * scal("t", 1.0)
            gmax = osp._seg_max(z, ei[1], N)
            e = (z - gmax.index_select(0, ei[1])).exp()
            w = e / osp._seg_sum(e, ei[1], N).index_select(0, ei[1])
        m = osp._seg_sum(msg * w, ei[1], N)
        if ref.aggr == "softmax_sum":
            m = torch.pow(osp.in_degree(ei[1], N, torch.float64).unsqueeze(1), torch.sigmoid(ref.y)) * m
        h = xr + (osp.msg_norm(xr, m, ref.msg_norm.msg_scale) if ref.msg_norm is not None else m)
    (h * wgt.double()).sum().backward()

    mod = mod.cuda().train()
    xc = x.cuda().requires_grad_(True)
    scale = mod.msg_norm.msg_scale if mod.msg_norm is not None else None
    hc = mod.propagate(ei.cuda(), x=xc, msg_scale=scale, residual=True)
    torch.testing.assert_close(hc.detach().cpu(), h.detach().float(), rtol=1e-3, atol=1e-4)
    (hc * wgt.cuda()).sum().backward()

    def check(name, got, want):
        scale_ = want.abs().max().clamp_min(1e-6)
        err = (got.cpu().double() - want).abs().max() / scale_
        assert err < 3e-3, (name, float(err))
    check("x", xc.grad, xr.grad)
    for name in ("t", "p", "y"):
        pr = getattr(ref, name, None)
        if torch.is_tensor(pr) and pr.requires_grad:
            check(name, getattr(mod, name).grad, pr.grad)
    if ref.msg_norm is not None and ref.msg_norm.msg_scale.requires_grad:
        check("msg_scale", mod.msg_norm.msg_scale.grad, ref.msg_norm.msg_scale.grad)


def test_edge_attr_gradient_and_full_layer():
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    g = torch.Generator().manual_seed(3)
    N, E, C = 120, 900, 32
    ei = torch.randint(0, N, (2, E), generator=g)
    x, ea = torch.randn(N, C, generator=g), torch.randn(E, 7, generator=g)
    torch.manual_seed(2)
    mod = S.GENConv(C, 48, aggr="softmax", t=0.5, learn_t=True, encode_edge=True, edge_feat_dim=7, mlp_layers=2,
                    norm="layer")
    ref = __import__("copy").deepcopy(mod).double()
    xr, ear = x.double().requires_grad_(True), ea.double().requires_grad_(True)
    out_ref = ref.mlp(osp.genconv_pre_mlp(xr, ei, ref.edge_encoder(ear), "softmax", ref.t, 1.0, 0.0, None, ref.eps))
    out_ref.square().sum().backward()
    mod = mod.cuda()
    xc, eac = x.cuda().requires_grad_(True), ea.cuda().requires_grad_(True)
    out = mod(xc, ei.cuda(), eac)
    torch.testing.assert_close(out.detach().cpu(), out_ref.detach().float(), rtol=1e-3, atol=1e-4)
    out.square().sum().backward()
    for got, want in ((xc.grad, xr.grad), (eac.grad, ear.grad), (mod.t.grad, ref.t.grad),
                      (mod.edge_encoder.weight.grad, ref.edge_encoder.weight.grad),
                      (mod.mlp[0].weight.grad, ref.mlp[0].weight.grad)):
        err = (got.cpu().double() - want).abs().max() / want.abs().max().clamp_min(1e-6)
        assert err < 3e-3, float(err)
