"""GENConv - API of the reference's gcn_lib/sparse/torch_vertex.py:12-88."""
from torch import nn

from .torch_nn import MLP, BondEncoder
from .torch_message import GenMessagePassing, MsgNorm

__all__ = ["GENConv"]


class GENConv(GenMessagePassing):
    """GENeralized graph convolution (softmax / power-mean aggregation).

    forward (torch_vertex.py:62-76): edge encoder -> [message + aggregate + MsgNorm
    + residual: ONE fused kernel over the cached CSR graph] -> MLP (torch)."""

    def __init__(self, in_dim, emb_dim, aggr="softmax", t=1.0, learn_t=False, p=1.0, learn_p=False,
                 y=0.0, learn_y=False, msg_norm=False, learn_msg_scale=True, encode_edge=False,
                 bond_encoder=False, edge_feat_dim=None, norm="batch", mlp_layers=2, eps=1e-7):
        super().__init__(aggr=aggr, t=t, learn_t=learn_t, p=p, learn_p=learn_p, y=y, learn_y=learn_y)
        channels = [in_dim] + [in_dim * 2] * (mlp_layers - 1) + [emb_dim]
        self.mlp = MLP(channels=channels, norm=norm, last_lin=True)
        self.msg_encoder = nn.ReLU()
        self.eps = eps
        self.encode_edge = encode_edge
        self.bond_encoder = bond_encoder
        self.msg_norm = MsgNorm(learn_msg_scale=learn_msg_scale) if msg_norm else None
        if self.encode_edge:
            if self.bond_encoder:
                self.edge_encoder = BondEncoder(emb_dim=in_dim)
            else:
                self.edge_encoder = nn.Linear(edge_feat_dim, in_dim)

    def forward(self, x, edge_index, edge_attr=None):
        if self.encode_edge and edge_attr is not None:
            edge_emb = self.edge_encoder(edge_attr)
        else:
            edge_emb = edge_attr
        scale = self.msg_norm.msg_scale if self.msg_norm is not None else None
        h = self.propagate(edge_index, x=x, edge_attr=edge_emb, msg_scale=scale, residual=True)
        return self.mlp(h)

    def message(self, x_j, edge_attr=None):
        """torch_vertex.py:78-85 (reference formula; the kernel fuses it)."""
        msg = x_j + edge_attr if edge_attr is not None else x_j
        return self.msg_encoder(msg) + self.eps

    def update(self, aggr_out):
        return aggr_out
