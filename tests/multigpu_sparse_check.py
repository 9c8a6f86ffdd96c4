"""torchrun entry (one rank per GPU): node-partitioned GENConv forward vs the single-GPU
full-graph forward of the same layer.  Used by tests/test_multigpu_gpu.py."""
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
    x = torch.randn(N, C, generator=g)
    for aggr in ("softmax_sg", "power", "max"):
        torch.manual_seed(1)
        conv = S.GENConv(C, C, aggr=aggr, t=0.1, p=2.0, msg_norm=True, mlp_layers=1, norm="layer").to(dev).eval()
        with torch.no_grad():
            full = conv(x.to(dev), ei.to(dev))
        part = P.GraphPartition(ei, N, rank, world, device=dev).exchange_halo_lists()
        out = P.genconv_forward_partitioned(conv, x[part.lo:part.hi].to(dev).contiguous(), part)
        torch.testing.assert_close(out, full[part.lo:part.hi], rtol=1e-5, atol=1e-6)
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_SPARSE_OK world=%d halo_rows=%d" % (world, part.n_halo))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
