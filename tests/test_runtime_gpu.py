"""Runtime contract of the drop-in modules (SURVEY.md 8b 'Threading / devices'): kernels run on
the caller's current stream, hold no global mutable state, are safe under concurrent host threads
(torch.nn.DataParallel runs one per replica) and deterministic under torch.utils.checkpoint."""
import threading

import pytest
import torch
from torch.utils.checkpoint import checkpoint

pytestmark = pytest.mark.gpu


def _dense_case(seed):
    from deep_gcns_torch_b200.gcn_lib import dense as D
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    mod = D.DynConv2d(16, 24, 8, 2, "edge", "relu", "batch", True).cuda().eval()
    x = torch.randn(3, 16, 384, 1, generator=g).cuda()
    return mod, x


def _sparse_case(seed):
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    conv = S.GENConv(64, 64, aggr="softmax", t=0.5, learn_t=True, msg_norm=True, mlp_layers=1, norm="layer").cuda()
    x = torch.randn(2000, 64, generator=g).cuda()
    ei = torch.randint(0, 2000, (2, 30000), generator=g).cuda()
    return conv, x, ei


def test_side_stream_matches_default_stream():
    mod, x = _dense_case(0)
    conv, xs, ei = _sparse_case(0)
    with torch.no_grad():
        ref_d, ref_s = mod(x), conv(xs, ei)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        got_d, got_s = mod(x), conv(xs, ei)
    side.synchronize()
    assert torch.equal(got_d, ref_d) and torch.equal(got_s, ref_s)


def test_concurrent_host_threads():
    cases = [_dense_case(i) for i in range(4)]
    with torch.no_grad():
        refs = [m(x).clone() for m, x in cases]
    torch.cuda.synchronize()
    outs, errs = [None] * 4, []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(5):
                    outs[i] = cases[i][0](cases[i][1])
            s.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for o, r in zip(outs, refs):
        assert torch.equal(o, r)


def test_checkpoint_recompute_gives_identical_gradients():
    conv, x, ei = _sparse_case(1)
    x1 = x.clone().requires_grad_(True)
    conv(x1, ei).square().sum().backward()
    g_plain = [x1.grad.clone()] + [p.grad.clone() for p in conv.parameters() if p.grad is not None]
    conv.zero_grad()
    x2 = x.clone().requires_grad_(True)
    checkpoint(conv, x2, ei, use_reentrant=False).square().sum().backward()
    g_ckpt = [x2.grad] + [p.grad for p in conv.parameters() if p.grad is not None]
    assert len(g_plain) == len(g_ckpt)
    for a, b in zip(g_plain, g_ckpt):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_dense_training_step_updates_like_reference_semantics():
    """One SGD step through ResDynBlock2d: loss decreases on a fixed batch, running stats move."""
    from deep_gcns_torch_b200.gcn_lib import dense as D
    torch.manual_seed(0)
    blk = D.ResDynBlock2d(16, 8, 2, "edge", "relu", "batch", True, True, 0.5).cuda().train()
    x = torch.randn(4, 16, 256, 1, device="cuda")
    tgt = torch.randn(4, 16, 256, 1, device="cuda")
    opt = torch.optim.SGD(blk.parameters(), lr=0.05)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = (blk(x) - tgt).square().mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    bn = blk.body.gconv.nn[2]
    assert int(bn.num_batches_tracked) == 5 and float(bn.running_mean.abs().sum()) > 0
