"""Sparse nn helpers - API of the reference's gcn_lib/sparse/torch_nn.py (plain
torch: GEMM / norm layers stay library calls)."""
from torch import nn

__all__ = ["act_layer", "norm_layer", "MultiSeq", "MLP", "AtomEncoder", "BondEncoder",
           "get_atom_feature_dims", "get_bond_feature_dims"]


def get_atom_feature_dims():
    """Sizes of the 9 OGB atom feature vocabularies (reference: utils/data_util.py:314-325)."""
    return [119, 4, 12, 12, 10, 6, 6, 2, 2]


def get_bond_feature_dims():
    """Sizes of the 3 OGB bond feature vocabularies (reference: utils/data_util.py:342-347)."""
    return [5, 6, 2]


def act_layer(act_type, inplace=False, neg_slope=0.2, n_prelu=1):
    """torch_nn.py:9-21."""
    kind = act_type.lower()
    if kind == "relu":
        return nn.ReLU(inplace)
    if kind == "leakyrelu":
        return nn.LeakyReLU(neg_slope, inplace)
    if kind == "prelu":
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    raise NotImplementedError("activation layer [%s] is not found" % kind)


def norm_layer(norm_type, nc):
    """torch_nn.py:23-34."""
    kind = norm_type.lower()
    if kind == "batch":
        return nn.BatchNorm1d(nc, affine=True)
    if kind == "layer":
        return nn.LayerNorm(nc, elementwise_affine=True)
    if kind == "instance":
        return nn.InstanceNorm1d(nc, affine=False)
    raise NotImplementedError("normalization layer [%s] is not found" % kind)


class MultiSeq(nn.Sequential):
    """torch_nn.py:37-47: Sequential whose stages may take / return tuples."""

    def forward(self, *inputs):
        for stage in self._modules.values():
            inputs = stage(*inputs) if type(inputs) == tuple else stage(inputs)
        return inputs


class MLP(nn.Sequential):
    """torch_nn.py:50-71: Linear -> norm -> act -> dropout per layer, the last layer
    bare when `last_lin`."""

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0., last_lin=False):
        layers = []
        last = len(channels) - 1
        for i in range(1, len(channels)):
            layers.append(nn.Linear(channels[i - 1], channels[i], bias))
            if i == last and last_lin:
                continue
            if norm is not None and norm.lower() != "none":
                layers.append(norm_layer(norm, channels[i]))
            if act is not None and act.lower() != "none":
                layers.append(act_layer(act))
            if drop > 0:
                layers.append(nn.Dropout2d(drop))
        self.m = layers
        super().__init__(*layers)


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
        super().__init__(get_atom_feature_dims(), emb_dim, "atom_embedding_list")


class BondEncoder(_SumOfEmbeddings):
    """torch_nn.py:95-113 (state_dict keys bond_embedding_list.<i>.weight)."""

    def __init__(self, emb_dim):
        super().__init__(get_bond_feature_dims(), emb_dim, "bond_embedding_list")
