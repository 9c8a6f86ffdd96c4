"""Dense graph construction - API of the reference's gcn_lib/dense/torch_edge.py,
computed by the fused distance/selection kernels (no (B,N,N) matrix, no topk)."""
import torch
from torch import nn

from ... import _native

__all__ = ["DenseDilated", "DenseDilatedKnnGraph", "DilatedKnnGraph", "dense_knn_matrix", "pairwise_distance"]


def _stochastic_columns(k, dilation, stochastic, epsilon, training):
    """Host side of torch_edge.py:19-29.  None = regular ranks 0, d, 2d, ...; else the
    k ranks to keep.  Consumes the CPU generator exactly like the reference: one
    `torch.rand(1)` per call whenever `stochastic` (it is evaluated before
    `and self.training`), `randperm(k*d)` only when the branch is taken."""
    if stochastic:
        if torch.rand(1) < epsilon and training:
            return torch.randperm(k * dilation)[:k].tolist()
    return None


class DenseDilated(nn.Module):
    """torch_edge.py:6-29: pick the dilated columns of a (2,B,N,K) neighbour list."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k

    def forward(self, edge_index):
        cols = _stochastic_columns(self.k, self.dilation, self.stochastic, self.epsilon, self.training)
        if cols is None:
            return edge_index[:, :, :, ::self.dilation]
        return edge_index[:, :, :, torch.as_tensor(cols, device=edge_index.device)]


def pairwise_distance(x):
    """torch_edge.py:32-42, for API completeness only ((B,N,C) -> (B,N,N)); the
    graph builders below never call it."""
    inner = -2 * torch.matmul(x, x.transpose(2, 1))
    sq = torch.sum(x * x, dim=-1, keepdim=True)
    return sq + inner + sq.transpose(2, 1)


def dense_knn_matrix(x, k=16):
    """torch_edge.py:45-58: x (B,C,N,1) -> int64 (2,B,N,k), self included, ascending."""
    with torch.no_grad():
        edge_index, _ = _native.knn_graph(x.detach(), k, 1)
    return edge_index


class DenseDilatedKnnGraph(nn.Module):
    """torch_edge.py:61-76.  The dilation is applied inside the selection kernel: only
    the k surviving ranks of the k*dilation sorted neighbours are written."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)
        self.knn = dense_knn_matrix

    def columns(self):
        return _stochastic_columns(self.k, self.dilation, self.stochastic, self.epsilon, self.training)

    def forward(self, x):
        with torch.no_grad():
            edge_index, _ = _native.knn_graph(x.detach(), self.k, self.dilation, cols=self.columns())
        return edge_index


class DilatedKnnGraph(nn.Module):
    """torch_edge.py:79-101: the reference loops over the batch calling
    torch_cluster.knn_graph (self excluded); here one launch with the self mask on."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)

    def forward(self, x):
        cols = _stochastic_columns(self.k, self.dilation, self.stochastic, self.epsilon, self.training)
        with torch.no_grad():
            edge_index, _ = _native.knn_graph(x.detach(), self.k, self.dilation, cols=cols, exclude_self=True)
        return edge_index
