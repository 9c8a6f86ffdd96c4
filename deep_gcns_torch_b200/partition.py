"""Node-partitioned GENConv over several GPUs (BASELINE config 5).

New functionality - the reference has no multi-GPU sparse path (its products training drops
cross-partition edges, utils/data_util.py:43-61); parity oracle = the single-device full-graph
forward.  Destination nodes (CSR rows) are split into `world` contiguous ranges.  Rank r owns
x rows [lo_r, hi_r) and every edge whose TARGET lies in its range; sources outside the range
are its halo.  Per layer:

    pack   send_rows -> dgcn_gather_rows            (rows other ranks need from me)
    comm   all_to_all_single (NCCL over NVLink; gloo on CPU in the host-logic tests)
    fuse   dgcn_genconv_aggregate(x_src = [local rows | halo rows], x_dst = local rows)

The exchange is the only collective; BatchNorm1d / MLP are row-local.  One process per GPU.
"""
import torch
import torch.distributed as dist


def row_ranges(num_nodes, world):
    """Contiguous, balanced row ranges [(lo, hi)] * world."""
    base, rem = divmod(num_nodes, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class GraphPartition:
    """Everything rank `rank` needs about its share of the graph (index tensors on `device`)."""

    def __init__(self, edge_index, num_nodes, rank, world, device=None):
        device = device if device is not None else edge_index.device
        ei = edge_index.cpu()
        self.rank, self.world, self.num_nodes = rank, world, num_nodes
        self.ranges = row_ranges(num_nodes, world)
        lo, hi = self.ranges[rank]
        self.lo, self.hi, self.n_local = lo, hi, hi - lo
        bounds = torch.tensor([r[0] for r in self.ranges] + [num_nodes])
        mine = (ei[1] >= lo) & (ei[1] < hi)
        self.edge_ids = mine.nonzero(as_tuple=True)[0]              # positions in the global edge list
        src, dst = ei[0][mine], ei[1][mine] - lo
        remote = (src < lo) | (src >= hi)
        halo_nodes = torch.unique(src[remote])                      # sorted global ids => grouped by owner
        owner = torch.bucketize(halo_nodes, bounds, right=True) - 1
        self.recv_counts = [int((owner == r).sum()) for r in range(world)]
        self.n_halo = int(halo_nodes.numel())
        # local numbering of sources: own rows first, then halo rows in the order they arrive
        remap = torch.full((num_nodes,), -1, dtype=torch.long)
        remap[lo:hi] = torch.arange(self.n_local)
        remap[halo_nodes] = self.n_local + torch.arange(self.n_halo)
        self.local_edge_index = torch.stack((remap[src], dst)).to(device)
        self.halo_nodes = halo_nodes                                # global ids I need, grouped by owner rank
        self.device = device
        self.send_rows = None                                       # filled by exchange_halo_lists()
        self.send_counts = None
        self._csr = None

    def exchange_halo_lists(self, group=None):
        """Tell every owner which of its rows I need; learn which of mine the others need."""
        world = self.world
        want = [self.halo_nodes[sum(self.recv_counts[:r]):sum(self.recv_counts[:r + 1])] - self.ranges[r][0]
                for r in range(world)]
        counts = torch.tensor(self.recv_counts, dtype=torch.long)
        comm_dev = self.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
        theirs = torch.empty(world, dtype=torch.long, device=comm_dev)
        dist.all_to_all_single(theirs, counts.to(comm_dev), group=group)
        self.send_counts = [int(v) for v in theirs.cpu()]
        recv = torch.empty(sum(self.send_counts), dtype=torch.long, device=comm_dev)
        dist.all_to_all_single(recv, torch.cat(want).to(comm_dev), output_split_sizes=self.send_counts,
                               input_split_sizes=self.recv_counts, group=group)
        self.send_rows = recv.to(self.device, torch.int32)          # local row ids, grouped by destination rank
        return self

    def csr(self):
        if self._csr is None:
            from . import _native
            self._csr = _native.csr_build(self.local_edge_index, self.n_local)
        return self._csr


def halo_exchange(x_local, part, gather=None, group=None):
    """[local rows | halo rows]: packs the rows the peers asked for and swaps them all-to-all."""
    if gather is None:
        from . import _native
        gather = _native.gather_rows
    send = gather(x_local, part.send_rows)
    recv = torch.empty((part.n_halo, x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
    dist.all_to_all_single(recv, send, output_split_sizes=part.recv_counts, input_split_sizes=part.send_counts,
                           group=group)
    return torch.cat((x_local, recv), 0)


def genconv_forward_partitioned(conv, x_local, part, edge_attr_local=None, group=None):
    """GENConv.forward (gcn_lib/sparse/torch_vertex.py:62-76) on this rank's rows: halo exchange,
    fused aggregate over the local CSR, row-local MLP.  Inference path (no autograd)."""
    from . import _native
    with torch.no_grad():
        x_src = halo_exchange(x_local, part, group=group)
        t, p, y = conv._scalars()
        scale = conv.msg_norm.msg_scale if conv.msg_norm is not None else None
        prm, _keep = _native.genconv_params(conv._check_aggr(), t, p, y, conv.eps, scale, add_residual=True)
        ea = edge_attr_local
        if ea is not None and conv.encode_edge:
            ea = conv.edge_encoder(ea)
        h = _native.genconv_aggregate(x_src, x_local, part.csr(), prm, ea)
        return conv.mlp(h)
