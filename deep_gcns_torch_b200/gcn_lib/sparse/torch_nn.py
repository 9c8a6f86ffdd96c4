"""Sparse nn helpers - API of the reference's gcn_lib/sparse/torch_nn.py (plain
torch: GEMM / norm layers stay library calls)."""
from torch import nn

__all__ = ["act_layer", "norm_layer", "MultiSeq", "MLP", "AtomEncoder", "BondEncoder",
           "get_atom_feature_dims", "get_bond_feature_dims"]

_ATOM_VOCAB = (119, 4, 12, 12, 10, 6, 6, 2, 2)     # OGB atom features (reference: utils/data_util.py:314-325)
_BOND_VOCAB = (5, 6, 2)                            # OGB bond features (reference: utils/data_util.py:342-347)

# name -> factory(inplace, negative slope, PReLU parameter count)          (torch_nn.py:9-21)
_ACTIVATIONS = {
    "relu": lambda inplace, slope, n: nn.ReLU(inplace),
    "leakyrelu": lambda inplace, slope, n: nn.LeakyReLU(slope, inplace),
    "prelu": lambda inplace, slope, n: nn.PReLU(num_parameters=n, init=slope),
}
# name -> factory(channels)                                                 (torch_nn.py:23-34)
_NORMS = {
    "batch": lambda nc: nn.BatchNorm1d(nc, affine=True),
    "layer": lambda nc: nn.LayerNorm(nc, elementwise_affine=True),
    "instance": lambda nc: nn.InstanceNorm1d(nc, affine=False),
}


def get_atom_feature_dims():
    return list(_ATOM_VOCAB)


def get_bond_feature_dims():
    return list(_BOND_VOCAB)


def _lookup(table, name, what):
    try:
        return table[name.lower()]
    except KeyError:
        raise NotImplementedError("%s layer [%s] is not found" % (what, name.lower())) from None


def act_layer(act_type, inplace=False, neg_slope=0.2, n_prelu=1):
    return _lookup(_ACTIVATIONS, act_type, "activation")(inplace, neg_slope, n_prelu)


def norm_layer(norm_type, nc):
    return _lookup(_NORMS, norm_type, "normalization")(nc)


class MultiSeq(nn.Sequential):
    """torch_nn.py:37-47: Sequential whose stages may take / return tuples."""

    def forward(self, *inputs):
        for stage in self._modules.values():
            inputs = stage(*inputs) if type(inputs) == tuple else stage(inputs)
        return inputs


def _named(option):
    return option is not None and option.lower() != "none"


class MLP(nn.Sequential):
    """torch_nn.py:50-71: Linear -> norm -> act -> dropout per layer, the last layer
    bare when `last_lin`."""

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0., last_lin=False):
        stack = []
        widths = list(zip(channels[:-1], channels[1:]))
        for pos, (c_in, c_out) in enumerate(widths):
            stack.append(nn.Linear(c_in, c_out, bias))
            if last_lin and pos == len(widths) - 1:
                break
            if _named(norm):
                stack.append(norm_layer(norm, c_out))
            if _named(act):
                stack.append(act_layer(act))
            if drop > 0:
                stack.append(nn.Dropout2d(drop))
        self.m = stack
        super().__init__(*stack)


class _SumOfEmbeddings(nn.Module):
    def __init__(self, dims, emb_dim, list_name):
        super().__init__()
        tables = nn.ModuleList()
        for dim in dims:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            tables.append(emb)
        setattr(self, list_name, tables)
        self._list_name = list_name

    def forward(self, feats):
        tables = getattr(self, self._list_name)
        total = 0
        for i in range(feats.shape[1]):
            total = total + tables[i](feats[:, i])
        return total


class AtomEncoder(_SumOfEmbeddings):
    """torch_nn.py:74-92 (state_dict keys atom_embedding_list.<i>.weight)."""

    def __init__(self, emb_dim):
        super().__init__(_ATOM_VOCAB, emb_dim, "atom_embedding_list")


class BondEncoder(_SumOfEmbeddings):
    """torch_nn.py:95-113 (state_dict keys bond_embedding_list.<i>.weight)."""

    def __init__(self, emb_dim):
        super().__init__(_BOND_VOCAB, emb_dim, "bond_embedding_list")
