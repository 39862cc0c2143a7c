"""Plain-aggregation GNNs of the reference (models/backbones.py:440-498 GraphSAGE, :246-300 GCNNet) on
the library's CSR SpMM.  GraphSAGE is what ``--no_dtc`` trains (main_graph_knowledge_transfer.py:414-417).
Layer parameter names follow PyG's SAGEConv / GCNConv so reference state_dicts load unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..data import add_self_loops, remove_self_loops
from .KTGNN import NodeLinear


def _linear(x, weight):
    """x W^T on the library's row-panel GEMM when the shape is one it takes (fp32 CUDA, widths <= 256), else ATen."""
    return ops.linear(x, weight, None) if ops.linear_supported(x, weight) else F.linear(x, weight)


class SAGEConv(nn.Module):
    """lin_l(mean_{j->i} x_j) + lin_r(x_i)  (PyG SAGEConv, aggr='mean'; lin_l has the bias)."""

    def __init__(self, in_channels, out_channels, root_weight=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.root_weight = in_channels, out_channels, root_weight
        self.lin_l = NodeLinear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = NodeLinear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        if self.root_weight:
            self.lin_r.reset_parameters()

    def forward(self, x, edge_index):
        graph = edge_index if isinstance(edge_index, ops.CSRGraph) else ops.cached_graph(edge_index, x.shape[0])
        if self.in_channels > self.out_channels:
            # mean aggregation commutes with the linear map: transform first, gather the narrower rows
            # (e.g. 1685 -> 64 one-hot inputs of the fb graphs: 26x less gather traffic)
            out = ops.spmm(graph, _linear(x, self.lin_l.weight), "mean")
            if self.lin_l.bias is not None:
                out = out + self.lin_l.bias          # lin_l(mean) = W mean + b also for rows without in-edges
        else:
            out = self.lin_l(ops.spmm(graph, x, "mean"))
        if self.root_weight:
            out = out + self.lin_r(x)
        return out


class GCNConv(nn.Module):
    """D^-1/2 (A + I) D^-1/2 X W + b  (PyG GCNConv with gcn_norm, add_self_loops=True)."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = NodeLinear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self._key, self._graph, self._dis = None, None, None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.lin.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def _norm(self, edge_index, n):
        key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, n)
        if self._key != key:
            ei = add_self_loops(remove_self_loops(edge_index), n)
            self._graph = ops.CSRGraph(ei, n)
            dis = self._graph.deg.pow(-0.5)
            self._dis = dis.masked_fill(torch.isinf(dis), 0.0).contiguous()
            self._key = key
        return self._graph, self._dis

    def forward(self, x, edge_index):
        graph, dis = self._norm(edge_index, x.shape[0])
        # w_ji = d_j^-1/2 d_i^-1/2 factored into a gather-side and a row-side scale: no per-edge weights in HBM
        out = ops.spmm(graph, self.lin(x), "sum", None, dis, dis)
        return out if self.bias is None else out + self.bias


def _stack(conv, dataset, layer_num, hidden, **kw):
    convs = nn.ModuleList()
    if layer_num == 1:
        convs.append(conv(dataset.num_features, dataset.num_classes, **kw))
    else:
        for num in range(layer_num):
            cin = dataset.num_features if num == 0 else hidden
            cout = dataset.num_classes if num == layer_num - 1 else hidden
            convs.append(conv(cin, cout, **kw))
    return convs


class _Stacked(nn.Module):
    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def _run(self, data, upto):
        x, edge_index = data.x, data.edge_index
        graph = self._graph(edge_index, x.shape[0])
        for ind, conv in enumerate(self.convs[:upto]):
            x = conv(x, graph)
            if ind != len(self.convs) - 1:
                x = F.dropout(F.relu(x), p=0.5, training=self.training)
        return x

    def _graph(self, edge_index, n):
        return edge_index

    def forward(self, data):
        return F.log_softmax(self._run(data, len(self.convs)), dim=1)

    def get_logits(self, data, layer_num=1):
        return self._run(data, len(self.convs))

    def get_emb(self, data, layer_num=1):
        return self._run(data, len(self.convs) - 1)


class GraphSAGE(_Stacked):
    def __init__(self, dataset, layer_num=2, hidden=16, root_weight=True):
        super().__init__()
        self.convs = _stack(SAGEConv, dataset, layer_num, hidden, root_weight=root_weight)

    def _graph(self, edge_index, n):
        # the reference builds SparseTensor(row=dst, col=src) on every forward (models/backbones.py:464);
        # here the CSR is built once per edge_index and shared by all layers
        return ops.cached_graph(edge_index, n)


class GCNNet(_Stacked):
    def __init__(self, dataset, layer_num=2, hidden=16):
        super().__init__()
        self.convs = _stack(GCNConv, dataset, layer_num, hidden)
