"""World-size-2 gloo test of the node-partition / halo-exchange host logic (no GPU): the
partitioned aggregation, with the oracle standing in for the CUDA kernel, must reproduce the
single-device result exactly."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sparse as osp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deep_gcns_torch_b200 import partition as P
    g = torch.Generator().manual_seed(0)
    N, E, C = 157, 1500, 12
    ei = torch.randint(0, N, (2, E), generator=g)
    x = torch.randn(N, C, generator=g)
    part = P.GraphPartition(ei, N, rank, world, device=torch.device("cpu")).exchange_halo_lists()
    lo, hi = part.lo, part.hi
    x_src = P.halo_exchange(x[lo:hi].contiguous(), part,
                            gather=lambda t, rows: t.index_select(0, rows.long()))
    assert x_src.shape[0] == part.n_local + part.n_halo
    assert torch.equal(x_src[part.n_local:], x[part.halo_nodes])           # halo rows arrived in order
    # every local edge points at the right source row
    src_global = ei[0][part.edge_ids]
    assert torch.equal(x_src[part.local_edge_index[0]], x[src_global])
    h = x[lo:hi] + osp.aggregate(osp.message(x_src, part.local_edge_index), part.local_edge_index[1],
                                 part.n_local, "softmax", 0.3)
    full = x + osp.aggregate(osp.message(x, ei), ei[1], N, "softmax", 0.3)
    torch.testing.assert_close(h, full[lo:hi], rtol=1e-6, atol=1e-6)
    out[rank] = (part.n_halo, sum(part.send_counts))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    # what rank 0 receives is what rank 1 sends and vice versa
    assert out[0][0] == out[1][1] and out[1][0] == out[0][1]


def test_row_ranges_cover():
    from deep_gcns_torch_b200.partition import row_ranges
    for n, w in ((10, 3), (8, 8), (5, 2), (2449029, 8)):
        r = row_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
