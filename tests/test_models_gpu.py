"""Model-level parity (BASELINE configs 2, 3, 4 in small): the drop-in classes assembled like the reference's
own models (bench_models.py restates examples/*/architecture.py / model.py with the same attribute names)
load the state_dict of the REFERENCE model and must reproduce its output.  Goldens:
tests/golden/gen_golden_models.py (reference model sources executed on the unmodified gcn_lib)."""
import os
import sys

import pytest
import torch

import golden_util as gu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


def _load_strict(model, sd):
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.cuda().eval()


def _frac_bad(got, ref):
    return float(((got - ref).abs() > ATOL + RTOL * ref.abs()).float().mean())


@pytest.fixture(autouse=True)
def _ieee_fp32_convolutions():
    """The model tails (fusion / prediction BasicConv heads) are plain torch Conv2d modules, exactly the reference's;
    torch runs cuDNN convolutions in TF32 by default (1e-3 relative), the CPU golden is fp32 - compare like with like."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _blockwise(model, c, tail):
    """Every stage of the model on the REFERENCE's own stage input (no error accumulation): graph adjudicated
    as a set per (cloud, point) row, features compared on the rows whose neighbour set agrees (a row differs
    only at an fp32 near-tie, tests/test_dense_gpu.py adjudicates those against fp64); then the tail (fusion /
    pooling / prediction, plain torch modules) on the reference's features; then the end-to-end output, where
    one near-tie flip anywhere reaches every logit through the global pooling - bounded loosely."""
    inputs = c.ins["inputs"].cuda()
    gold_feats = [c.outs["feat%d" % i] for i in range(len(model.backbone) + 1)]
    gold_graphs = [c.outs["graph%d" % i].long() for i in range(len(model.backbone) + 1)]
    with torch.no_grad():
        ei = model.knn(inputs[:, 0:3])
        same = (ei[0].cpu().sort(-1).values == gold_graphs[0].sort(-1).values).all(-1)
        assert same.float().mean() >= 0.99
        f0 = model.head(inputs, ei).cpu()
        mask = same.unsqueeze(1).unsqueeze(-1).expand_as(f0)
        torch.testing.assert_close(f0[mask], gold_feats[0][mask], rtol=RTOL, atol=ATOL)
        # the head on the reference's graph: no masking needed
        gold_ei = torch.stack((gold_graphs[0], torch.arange(gold_graphs[0].shape[1]).view(1, -1, 1).expand_as(gold_graphs[0])))
        torch.testing.assert_close(model.head(inputs, gold_ei.cuda()).cpu(), gold_feats[0], rtol=RTOL, atol=ATOL)
        for i, blk in enumerate(model.backbone):
            x_in = gold_feats[i].cuda()
            g = blk.body.dilated_knn_graph(x_in)[0].cpu()
            same = (g.sort(-1).values == gold_graphs[i + 1].sort(-1).values).all(-1)
            assert same.float().mean() >= 0.99, (i, float(same.float().mean()))
            out = blk(x_in).cpu()
            mask = same.unsqueeze(1).unsqueeze(-1).expand_as(out)
            torch.testing.assert_close(out[mask], gold_feats[i + 1][mask], rtol=RTOL, atol=ATOL)
        y_tail = tail(model, [f.cuda() for f in gold_feats]).cpu()
        torch.testing.assert_close(y_tail, c.outs["y"], rtol=RTOL, atol=ATOL)
        y = model(inputs).cpu()
    err = (y - c.outs["y"]).abs()
    scale = float(c.outs["y"].abs().max())
    assert float(err.max()) <= 0.05 * scale, (float(err.max()), scale)
    assert float((err <= ATOL + RTOL * c.outs["y"].abs()).float().mean()) >= 0.5 or float(err.max()) <= 0.01 * scale


def test_resgcn_4_blocks_matches_reference_model():
    """DenseDeepGCN (sem_seg_dense): kNN head on inputs[:, 0:3], EdgeConv head, 3 ResDynBlock2d with dilation
    1..3 (stochastic dilation in eval = regular dilation + one host RNG draw per layer), fusion, prediction."""
    from bench_models import ResGCN28
    from deep_gcns_torch_b200.gcn_lib import dense as D
    c = gu.load("model_resgcn4")
    m = c.meta
    model = _load_strict(ResGCN28(D, m["in_channels"], m["n_classes"], m["k"], m["n_filters"], m["n_blocks"]), c.sd)

    def tail(mod, feats):
        feats = torch.cat(feats, dim=1)
        fusion = torch.max_pool2d(mod.fusion_block(feats), kernel_size=[feats.shape[2], feats.shape[3]])
        fusion = torch.repeat_interleave(fusion, repeats=feats.shape[2], dim=2)
        return mod.prediction(torch.cat((fusion, feats), dim=1)).squeeze(-1)
    torch.manual_seed(5)
    _blockwise(model, c, tail)


def test_mrgcn_4_blocks_matches_reference_model():
    """DeepGCN (modelnet_cls): DilatedKnnGraph head (self excluded), MRConv head, 3 ResDynBlock2d('mr') with
    dilation 1..3, fusion + max/avg pooling + prediction."""
    import torch.nn.functional as F
    from bench_models import MRGCN28
    from deep_gcns_torch_b200.gcn_lib import dense as D
    c = gu.load("model_mrgcn4")
    m = c.meta
    model = _load_strict(MRGCN28(D, m["in_channels"], m["n_classes"], m["k"], m["n_filters"], m["n_blocks"],
                                 m["emb_dims"]), c.sd)

    def tail(mod, feats):
        fusion = mod.fusion_block(torch.cat(feats, dim=1))
        x1, x2 = F.adaptive_max_pool2d(fusion, 1), F.adaptive_avg_pool2d(fusion, 1)
        return mod.prediction(torch.cat((x1, x2), dim=1)).squeeze(-1).squeeze(-1)
    _blockwise(model, c, tail)


def test_deepergcn_8_layers_res_plus_matches_reference_model():
    from bench_models import DeeperGCN
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    c = gu.load("model_deepergcn8")
    m = c.meta
    model = _load_strict(DeeperGCN(S, m["num_layers"], m["hidden_channels"], m["in_channels"], m["num_tasks"]), c.sd)
    x, ei = c.ins["x"].cuda(), c.ins["edge_index"].long().cuda()
    with torch.no_grad():
        y = model(x, ei)
        y_fused = model.forward_fused(x, ei)
    torch.testing.assert_close(y.cpu(), c.outs["y"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(y_fused.cpu(), c.outs["y"], rtol=RTOL, atol=ATOL)     # opt-in fused res+ blocks


class _Coupling(torch.nn.Module):
    """eff_gcn_modules/rev/memgcn.py:9-52 restated (test infrastructure): chunk channels into groups, every
    group's GENConv sees non-contiguous channel slices of x and of the edge features."""

    def __init__(self, fms, group):
        super().__init__()
        self.Fms, self.group = fms, group

    def forward(self, x, edge_index, *args):
        xs = torch.chunk(x, self.group, dim=-1)
        chunks = list(zip(*[torch.chunk(a, self.group, dim=-1) for a in args]))
        y_in = sum(xs[1:])
        ys = []
        for i in range(self.group):
            y_in = xs[i] + self.Fms[i](y_in, edge_index, *chunks[i])
            ys.append(y_in)
        return torch.cat(ys, dim=-1)

    def inverse(self, y, edge_index, *args):
        ys = torch.chunk(y, self.group, dim=-1)
        chunks = list(zip(*[torch.chunk(a, self.group, dim=-1) for a in args]))
        xs = []
        for i in range(self.group - 1, -1, -1):
            y_in = ys[i - 1] if i != 0 else sum(xs)
            xs.append(ys[i] - self.Fms[i](y_in, edge_index, *chunks[i]))
        return torch.cat(xs[::-1], dim=-1)


def test_genconv_under_group_additive_coupling():
    """RevGNN: GENConv called on chunked (strided, non-contiguous) inputs and edge features, forward and
    inverse; the reconstruction must also close on our own outputs."""
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    c = gu.load("model_revgnn")
    m = c.meta
    cg = m["C"] // m["group"]
    fms = torch.nn.ModuleList(S.GENConv(cg, cg, aggr=m["aggr"], t=m["t"], learn_t=m["learn_t"], msg_norm=m["msg_norm"],
                                        learn_msg_scale=m["learn_msg_scale"], norm=m["norm"], mlp_layers=m["mlp_layers"])
                              for _ in range(m["group"]))
    model = _load_strict(_Coupling(fms, m["group"]), c.sd)
    x, ei, ea = c.ins["x"].cuda(), c.ins["edge_index"].long().cuda(), c.ins["edge_attr"].cuda()
    with torch.no_grad():
        y = model(x, ei, ea)
        back_gold = model.inverse(c.outs["y"].cuda(), ei, ea)
        back_own = model.inverse(y, ei, ea)
    torch.testing.assert_close(y.cpu(), c.outs["y"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(back_gold.cpu(), c.outs["x_back"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(back_own, x, rtol=RTOL, atol=1e-3)
    # and through autograd (the reversible wrapper differentiates through the module)
    model.train()
    xg = x.clone().requires_grad_(True)
    model(xg, ei, ea).square().sum().backward()
    assert torch.isfinite(xg.grad).all() and float(xg.grad.abs().sum()) > 0
