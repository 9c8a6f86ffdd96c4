"""torchrun entry (one rank per GPU): node-partitioned GENConv vs the single-GPU full-graph result of the
same layer - forward on the overlapped persistent-buffer path (with and without the fused norm -> relu
pre-activation), and forward + backward on the autograd path (reverse halo exchange).
Used by tests/test_multigpu_gpu.py."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from deep_gcns_torch_b200 import partition as P
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    g = torch.Generator().manual_seed(0)
    N, E, C = 20011, 300000, 128
    ei = torch.randint(0, N, (2, E), generator=g)
    ei[1, :5000] = 17                                  # one hub row (segmented CTA path) on rank 0
    x = torch.randn(N, C, generator=g)
    part = P.GraphPartition(ei.to(dev), N, rank, world, device=dev).exchange_halo_lists(P.high_priority_group())
    lo, hi = part.lo, part.hi
    for aggr in ("softmax_sg", "power", "max"):
        torch.manual_seed(1)
        conv = S.GENConv(C, C, aggr=aggr, t=0.1, p=2.0, msg_norm=True, mlp_layers=1, norm="layer").to(dev).eval()
        with torch.no_grad():
            full = conv(x.to(dev), ei.to(dev))
            out = P.genconv_forward_partitioned(conv, x[lo:hi].to(dev).contiguous(), part)
        torch.testing.assert_close(out, full[lo:hi], rtol=1e-5, atol=1e-6)
        # fused pre-activation: the buffer holds raw h, the kernel reads relu(s * h + t)
        s = (torch.rand(C, generator=g) + 0.5).to(dev)
        t = torch.randn(C, generator=g).to(dev) * 0.1
        with torch.no_grad():
            z = torch.relu(x.to(dev) * s + t)
            full_pre = conv.propagate(ei.to(dev), x=z, msg_scale=conv.msg_norm.msg_scale, residual=True)
            part.local_rows(C).copy_(x[lo:hi].to(dev))
            got = P.aggregate_partitioned(conv, part, C, pre=(s, t, True))
            got_serial = P.aggregate_partitioned(conv, part, C, pre=(s, t, True), overlap=False)
        torch.testing.assert_close(got, full_pre[lo:hi], rtol=1e-5, atol=1e-6)
        assert torch.equal(got, got_serial)            # split launches change nothing
    # training path: gradients w.r.t. x and the parameters through the reverse exchange
    torch.manual_seed(2)
    conv = S.GENConv(C, C, aggr="softmax", t=0.5, learn_t=True, msg_norm=True, learn_msg_scale=True, mlp_layers=1,
                     norm="layer").to(dev).train()
    w = torch.randn(N, C, generator=g).to(dev)
    xf = x.to(dev).clone().requires_grad_(True)
    (conv(xf, ei.to(dev)) * w).sum().backward()
    ref_gx = xf.grad.clone()
    ref_gp = {n: p.grad.clone() for n, p in conv.named_parameters() if p.grad is not None}
    conv.zero_grad()
    xl = x[lo:hi].to(dev).clone().requires_grad_(True)
    (P.genconv_forward_partitioned(conv, xl, part) * w[lo:hi]).sum().backward()
    torch.testing.assert_close(xl.grad, ref_gx[lo:hi], rtol=2e-4, atol=2e-5)
    for n, p in conv.named_parameters():
        if p.grad is None:
            continue
        gsum = p.grad.clone()
        dist.all_reduce(gsum)                          # parameter gradients add up over the row partitions
        # sums of ~20k row products: compare at the scale of the gradient, not of its smallest entries
        scale = float(ref_gp[n].abs().max())
        torch.testing.assert_close(gsum, ref_gp[n], rtol=2e-3, atol=2e-4 * max(scale, 1.0),
                                   msg=lambda m, n=n: "%s: %s" % (n, m))
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_SPARSE_OK world=%d halo_rows=%d interior=%d boundary=%d" %
              (world, part.n_halo, part.interior_rows.numel(), part.boundary_rows.numel()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
