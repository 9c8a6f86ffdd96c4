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
    # persistent-buffer path: the layer input sits in the buffer's local region, the halo lands behind it
    part.local_rows(C).copy_(x[lo:hi])
    assert P.start_halo_exchange(part, C) is None                          # CPU: synchronous
    xbuf, _send = part.buffers(C)
    assert torch.equal(xbuf, x_src) and xbuf.data_ptr() == part.buffers(C)[0].data_ptr()
    # interior rows never read a halo row, boundary rows do; together they cover the rank's rows
    src_l, dst_l = part.local_edge_index
    reads_halo = torch.zeros(part.n_local, dtype=torch.bool)
    reads_halo[dst_l[src_l >= part.n_local]] = True
    assert torch.equal(part.interior_rows.long(), (~reads_halo).nonzero()[:, 0])
    assert torch.equal(part.boundary_rows.long(), reads_halo.nonzero()[:, 0])
    assert part.halo_bytes(C) == part.n_halo * C * 4
    # reverse exchange: d/dx of sum(w * [local | halo]) through HaloExchange equals the single-process gradient
    w = torch.randn(N, C, generator=g)
    xl = x[lo:hi].clone().requires_grad_(True)
    xs = P.HaloExchange.apply(xl, part, None)
    w_src = torch.cat((w[lo:hi], w[part.halo_nodes]))
    (xs * w_src).sum().backward()
    # every rank r' that needs my row i contributes w[i]; plus my own copy
    need = torch.zeros(N)
    for r in range(world):
        pr = P.GraphPartition(ei, N, r, world, device=torch.device("cpu"))
        need[pr.halo_nodes] += 1
    expect = w[lo:hi] * (1 + need[lo:hi]).unsqueeze(1)
    torch.testing.assert_close(xl.grad, expect, rtol=1e-6, atol=1e-6)
    out[rank] = (part.n_halo, sum(part.send_counts))
    dist.barrier()
    dist.destroy_process_group()


def _world1_worker(rank, world, port, out):
    """world = 1 and a block-diagonal graph: empty halo lists everywhere must be legal."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deep_gcns_torch_b200 import partition as P
    g = torch.Generator().manual_seed(1)
    N, C = 40, 6
    lo, hi = P.row_ranges(N, world)[rank]
    n = hi - lo
    ei = torch.stack((torch.randint(lo, hi, (200,), generator=g), torch.randint(lo, hi, (200,), generator=g)))
    part = P.GraphPartition.from_local_edges(ei[0], ei[1], N, rank, world).exchange_halo_lists()
    assert part.n_halo == 0 and part.send_rows.numel() == 0 and part.boundary_rows.numel() == 0
    x = torch.randn(n, C, generator=g)
    assert torch.equal(P.halo_exchange(x, part), x)
    part.local_rows(C).copy_(x)
    P.start_halo_exchange(part, C)
    assert torch.equal(part.buffers(C)[0], x)
    out[rank] = True
    dist.barrier()
    dist.destroy_process_group()


def test_empty_halo_world1_and_block_diagonal_world2():
    for world in (1, 2):
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_world1_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert len(out) == world


def test_bfs_order_recovers_locality():
    """A banded graph whose ids were shuffled: the Cuthill-McKee order brings the halo of a contiguous
    4-way split back from 'almost everything' to a few band widths."""
    from deep_gcns_torch_b200 import partition as P
    g = torch.Generator().manual_seed(0)
    N, band, world = 4000, 6, 4
    i = torch.arange(N).repeat_interleave(band)
    j = (i + torch.randint(1, band + 1, (N * band,), generator=g)).clamp(max=N - 1)
    ei = torch.cat((torch.stack((i, j)), torch.stack((j, i))), 1)
    shuffle = torch.randperm(N, generator=g)
    ei_shuffled = shuffle[ei]

    def total_halo(e):
        return sum(P.GraphPartition(e, N, r, world).n_halo for r in range(world))
    order = P.bfs_order(ei_shuffled, N)
    assert torch.equal(torch.sort(order).values, torch.arange(N))          # a permutation
    ei_new, perm = P.relabel(ei_shuffled, order)
    assert torch.equal(perm[order], torch.arange(N))
    before, after, ideal = total_halo(ei_shuffled), total_halo(ei_new), total_halo(ei)
    assert before > 0.5 * N * (world - 1) / world * world * 0.5            # shuffled: most remote nodes are halo
    assert after <= 4 * max(ideal, 2 * band * (world - 1)), (before, after, ideal)


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


def _dense_worker(rank, world, port, out):
    """Dense path: the cloud batch is sharded, no data-path collective.  With the oracle
    standing in for the CUDA layer, the all-gathered shard outputs must equal the full-batch result
    bit for bit (clouds never interact), and the bench's max-over-ranks time reduction must agree."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deep_gcns_torch_b200.partition import row_ranges
    from oracle import dense as od
    g = torch.Generator().manual_seed(0)
    B, C, N, k, d = 6, 8, 96, 5, 2
    x = torch.randn(B, C, N, 1, generator=g)
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(2 * C, 10, 1)
    p = {"weight": conv.weight.detach(), "bias": conv.bias.detach()}
    lo, hi = row_ranges(B, world)[rank]
    y_loc = od.dyn_conv(x[lo:hi], p, k, d, "edge", "relu", None)
    ei_loc = od.dilated_knn_graph(x[lo:hi], k, d)
    ys = [torch.empty_like(y_loc) for _ in range(world)]
    eis = [torch.empty_like(ei_loc) for _ in range(world)]
    dist.all_gather(ys, y_loc)
    dist.all_gather(eis, ei_loc.contiguous())
    assert torch.equal(torch.cat(ys, 0), od.dyn_conv(x, p, k, d, "edge", "relu", None))
    assert torch.equal(torch.cat(eis, 1), od.dilated_knn_graph(x, k, d))
    t = torch.tensor([1.0 + rank, 2.0 - rank], dtype=torch.float64)      # per-rank "timings"
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = t.tolist()
    dist.barrier()
    dist.destroy_process_group()


def test_dense_batch_sharding_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dense_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == out[1] == [2.0, 2.0]
