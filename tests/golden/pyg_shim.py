"""Minimal stand-ins for torch_geometric / torch_sparse / torch_scatter.

TEST INFRASTRUCTURE ONLY.  The reference (wendongbi/Bridged-GNN) is pure
Python on top of PyG, which is not installed in this image.  This file
restates the *published* semantics of the handful of PyG entry points the
reference's hot path calls, so that ``make_golden.py`` can import the
reference's own, unmodified ``models/KTGNN.py``, ``models/models.py``,
``models/backbones.py`` and ``main_bridged_graph.py`` from /root/reference and
record golden input/output vectors.  PyG version is unpinned upstream (no
requirements file); semantics below follow PyG 2.1-2.3, the era the
reference's __pycache__ (CPython 3.7) implies.

Nothing in the product imports this module.
"""
import inspect
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- utils
def _scatter_add(src, index, dim_size):
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def _scatter_max(src, index, dim_size):
    out = torch.full((dim_size,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """torch_geometric.utils.softmax (index form): exp(s - max_dst) / (sum_dst + 1e-16)."""
    assert ptr is None and dim == 0
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    src_max = _scatter_max(src.detach(), index, n)
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = _scatter_add(out, index, n) + 1e-16
    return out / out_sum.index_select(0, index)


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    assert edge_attr is None
    return torch.cat([edge_index, loop], dim=1), None


def add_remaining_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    mask = edge_index[0] != edge_index[1]
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        loop_attr = torch.full((n,), fill_value, dtype=edge_attr.dtype, device=edge_attr.device)
        inv = ~mask
        loop_attr[edge_index[0][inv]] = edge_attr[inv]
        edge_attr = torch.cat([edge_attr[mask], loop_attr], dim=0)
    return torch.cat([edge_index[:, mask], loop], dim=1), edge_attr


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((n,), dtype=dtype or torch.float32, device=index.device)
    return out.scatter_add_(0, index, torch.ones_like(index, dtype=out.dtype))


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    """torch_geometric.utils.coalesce: sort by row*N+col, drop duplicates."""
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    key = edge_index[0] * n + edge_index[1]
    key, perm = torch.sort(key)
    edge_index = edge_index[:, perm]
    mask = torch.ones_like(key, dtype=torch.bool)
    mask[1:] = key[1:] != key[:-1]
    assert edge_attr is None
    return edge_index[:, mask]


def to_undirected(edge_index, num_nodes=None):
    row, col = edge_index
    ei = torch.stack([torch.cat([row, col]), torch.cat([col, row])], dim=0)
    return coalesce(ei, num_nodes=num_nodes)


# --------------------------------------------------------------------------- nn
class Linear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear (state_dict keys weight/bias)."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer = weight_initializer
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            a = (6.0 / (self.weight.size(0) + self.weight.size(1))) ** 0.5
            nn.init.uniform_(self.weight, -a, a)
        else:
            nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if self.bias is not None:
            bound = 1.0 / (self.in_channels ** 0.5) if self.in_channels > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class SparseTensor:
    """torch_sparse.SparseTensor, COO only; enough for matmul(reduce=...)."""

    def __init__(self, row=None, col=None, value=None, sparse_sizes=None, **kw):
        self.row, self.col, self.value, self.sparse_sizes = row, col, value, sparse_sizes

    def set_value(self, value, layout=None):
        return SparseTensor(self.row, self.col, value, self.sparse_sizes)

    def to(self, *a, **k):
        return self

    def to_dense(self):
        n, m = self.sparse_sizes
        out = torch.zeros((n, m), dtype=self.value.dtype if self.value is not None else torch.float32)
        v = self.value if self.value is not None else torch.ones(self.row.numel())
        out.index_put_((self.row, self.col), v, accumulate=True)
        return out


def matmul(src, other, reduce="sum"):
    """torch_sparse.matmul(adj, x, reduce): out[row] = reduce_{(row,col)} value * x[col]."""
    msg = other.index_select(0, src.col)
    if src.value is not None:
        msg = msg * src.value.view(-1, 1).to(msg.dtype)
    n = src.sparse_sizes[0]
    out = _scatter_add(msg, src.row, n)
    if reduce in ("sum", "add"):
        return out
    if reduce == "mean":
        cnt = _scatter_add(torch.ones_like(src.row, dtype=other.dtype), src.row, n).clamp_(min=1)
        return out / cnt.view(-1, 1)
    raise NotImplementedError(reduce)


class MessagePassing(nn.Module):
    """torch_geometric.nn.conv.MessagePassing: gather (_j = edge_index[0], _i = edge_index[1]),
    message, aggregate by edge_index[1] (flow source_to_target)."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kw):
        super().__init__()
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        assert flow == "source_to_target" and node_dim == 0

    def propagate(self, edge_index, size=None, **kwargs):
        if isinstance(edge_index, SparseTensor):
            return self.message_and_aggregate(edge_index, kwargs["x"])
        params = list(inspect.signature(self.message).parameters)
        n = None
        args = {}
        for p in params:
            if p.endswith("_j") or p.endswith("_i"):
                data = kwargs[p[:-2]]
                if isinstance(data, (tuple, list)):
                    data = data[0] if p.endswith("_j") else data[1]
                n = data.shape[0]
                args[p] = data.index_select(0, edge_index[0] if p.endswith("_j") else edge_index[1])
            else:
                args[p] = kwargs[p]
        if n is None:
            x = kwargs["x"]
            n = (x[1] if isinstance(x, (tuple, list)) else x).shape[0]
        msg = self.message(**args)
        out = _scatter_add(msg, edge_index[1], n)
        if self.aggr == "add":
            return out
        if self.aggr == "mean":
            cnt = _scatter_add(torch.ones_like(edge_index[1], dtype=msg.dtype), edge_index[1], n).clamp_(min=1)
            return out / cnt.view(-1, 1)
        raise NotImplementedError(self.aggr)

    def message(self, x_j):
        return x_j


class SAGEConv(MessagePassing):
    """torch_geometric.nn.SAGEConv (aggr='mean'): lin_l(mean_j x_j) + lin_r(x_i); lin_l has the bias."""

    def __init__(self, in_channels, out_channels, normalize=False, root_weight=True, bias=True, **kw):
        super().__init__(aggr="mean")
        self.root_weight = root_weight
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        if self.root_weight:
            self.lin_r.reset_parameters()

    def forward(self, x, edge_index, size=None):
        out = self.propagate(edge_index, x=(x, x))
        out = self.lin_l(out)
        if self.root_weight:
            out = out + self.lin_r(x)
        return out

    def message_and_aggregate(self, adj_t, x):
        adj_t = adj_t.set_value(None)
        return matmul(adj_t, x[0], reduce=self.aggr)


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True, dtype=None):
    fill = 2.0 if improved else 1.0
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype or torch.float32, device=edge_index.device)
    if add_self_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = _scatter_add(edge_weight, col, num_nodes)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


class GCNConv(MessagePassing):
    """torch_geometric.nn.GCNConv: D^-1/2 (A+I) D^-1/2 (X W) + b."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=True,
                 normalize=True, bias=True, **kw):
        super().__init__(aggr="add")
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def reset_parameters(self):
        self.lin.reset_parameters()
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_weight=None):
        edge_index, w = gcn_norm(edge_index, edge_weight, x.size(0), dtype=x.dtype)
        x = self.lin(x)
        out = self.propagate(edge_index, x=x, edge_weight=w)
        if self.bias is not None:
            out = out + self.bias
        return out

    def message(self, x_j, edge_weight):
        return edge_weight.view(-1, 1) * x_j


class Data:
    """torch_geometric.data.Data: attribute bag."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.shape[0]

    @property
    def num_features(self):
        return self.x.shape[1]

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def coalesce(self):
        self.edge_index = coalesce(self.edge_index, num_nodes=self.num_nodes)
        return self


class ToUndirected:
    def __init__(self, reduce="add", merge=True):
        pass

    def __call__(self, data):
        data.edge_index = to_undirected(data.edge_index, data.num_nodes)
        return data


class _Unused(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the hot path; not provided by the shim")


def install():
    """Register the stand-in modules in sys.modules (idempotent)."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_bgnn_shim", False):
        return
    mk = types.ModuleType
    tg = mk("torch_geometric"); tg._bgnn_shim = True
    tg_nn = mk("torch_geometric.nn"); tg_conv = mk("torch_geometric.nn.conv")
    tg_dense = mk("torch_geometric.nn.dense"); tg_lin = mk("torch_geometric.nn.dense.linear")
    tg_utils = mk("torch_geometric.utils"); tg_typing = mk("torch_geometric.typing")
    tg_data = mk("torch_geometric.data"); tg_tf = mk("torch_geometric.transforms")
    tg_gcn = mk("torch_geometric.nn.conv.gcn_conv")
    ts = mk("torch_sparse"); ts_mm = mk("torch_sparse.matmul"); tsc = mk("torch_scatter")

    for name in ("SplineConv", "GATConv", "GATv2Conv", "GCN2Conv", "GENConv", "DeepGCNLayer", "APPNP",
                 "JumpingKnowledge", "GINConv"):
        setattr(tg_nn, name, type(name, (_Unused,), {}))
    tg_nn.MessagePassing = MessagePassing; tg_nn.SAGEConv = SAGEConv; tg_nn.GCNConv = GCNConv
    tg_nn.conv = tg_conv; tg_nn.dense = tg_dense
    tg_conv.MessagePassing = MessagePassing
    tg_conv.gat_conv = mk("torch_geometric.nn.conv.gat_conv")
    tg_conv.sage_conv = mk("torch_geometric.nn.conv.sage_conv")
    tg_conv.gcn_conv = tg_gcn
    tg_gcn.gcn_norm = gcn_norm
    tg_dense.linear = tg_lin; tg_lin.Linear = Linear
    for f in (softmax, remove_self_loops, add_self_loops, add_remaining_self_loops, degree, coalesce, to_undirected):
        setattr(tg_utils, f.__name__, f)
    from typing import Optional, Tuple, Union
    tg_typing.OptPairTensor = Tuple[torch.Tensor, Optional[torch.Tensor]]
    tg_typing.Adj = Union[torch.Tensor, SparseTensor]
    tg_typing.Size = Optional[Tuple[int, int]]
    tg_typing.NoneType = type(None)
    tg_typing.OptTensor = Optional[torch.Tensor]
    tg_data.Data = Data
    tg_data.InMemoryDataset = object
    tg_data.download_url = None
    tg_tf.ToUndirected = ToUndirected
    tg.nn, tg.utils, tg.typing, tg.data, tg.transforms = tg_nn, tg_utils, tg_typing, tg_data, tg_tf
    ts.SparseTensor = SparseTensor; ts.matmul = matmul
    ts.fill_diag = ts.sum = ts.mul = ts.set_diag = None
    sys.modules.update({
        "torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.nn.conv": tg_conv,
        "torch_geometric.nn.conv.gcn_conv": tg_gcn, "torch_geometric.nn.conv.gat_conv": tg_conv.gat_conv,
        "torch_geometric.nn.conv.sage_conv": tg_conv.sage_conv,
        "torch_geometric.nn.dense": tg_dense, "torch_geometric.nn.dense.linear": tg_lin,
        "torch_geometric.utils": tg_utils, "torch_geometric.typing": tg_typing,
        "torch_geometric.data": tg_data, "torch_geometric.transforms": tg_tf,
        "torch_sparse": ts, "torch_sparse.matmul": ts_mm, "torch_scatter": tsc,
    })
