"""Minimal graph container and edge-list utilities (torch_geometric is not a dependency).

``Data`` carries what the reference's hot path reads from a ``torch_geometric.data.Data``:
``x, edge_index, y, central_mask, train/val/test_mask``.  The edge-list helpers restate the PyG
calls the reference makes around the hot path (coalesce: main_bridged_graph.py:75,113,193;
ToUndirected: main_graph_knowledge_transfer.py:410-411; remove/add_self_loops: models/KTGNN.py:390-394)
on the device, using the library's radix-sort CSR builder for the sort/unique step.
"""
import sys
import types

import torch

from . import ops


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    @property
    def num_nodes(self):
        return self.x.shape[0]

    @property
    def num_features(self):
        return self.x.shape[1]

    def __repr__(self):
        parts = ["%s=%s" % (k, tuple(getattr(self, k).shape) if torch.is_tensor(getattr(self, k)) else getattr(self, k))
                 for k in self.keys()]
        return "Data(%s)" % ", ".join(parts)


def coalesce(edge_index, num_nodes=None):
    if edge_index.is_cuda:
        return ops.coalesce(edge_index, num_nodes)
    raise RuntimeError("coalesce runs on the device; move edge_index to CUDA first")


def to_undirected(edge_index, num_nodes):
    """ToUndirected(merge=True): add reversed edges, then coalesce."""
    both = torch.cat((edge_index, edge_index.flip(0)), dim=1)
    return coalesce(both, num_nodes)


def remove_self_loops(edge_index):
    return edge_index[:, edge_index[0] != edge_index[1]]


def add_self_loops(edge_index, num_nodes):
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat((edge_index, loop.unsqueeze(0).repeat(2, 1)), dim=1)


def load_pyg_dat(path):
    """Read a pickled ``torch_geometric.data.Data`` (the reference's ``data_bridged_graph/*.dat``,
    written at main_bridged_graph.py:317-320) without PyG: the pickle only needs attribute-bag classes
    under the PyG module names."""
    class _Bag:
        def __setstate__(self, s):
            self.__dict__.update(s)

    names = {"torch_geometric": [], "torch_geometric.data": [],
             "torch_geometric.data.data": ["Data", "DataEdgeAttr", "DataTensorAttr"],
             "torch_geometric.data.storage": ["GlobalStorage", "BaseStorage", "NodeStorage", "EdgeStorage"]}
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for mod, classes in names.items():
            if saved[mod] is None:
                m = types.ModuleType(mod)
                for c in classes:
                    setattr(m, c, type(c, (_Bag,), {}))
                sys.modules[mod] = m
        obj = torch.load(path, map_location="cpu", weights_only=False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
    store = obj.__dict__["_store"].__dict__["_mapping"]
    return Data(**{k: v for k, v in store.items()})


def save_graph(data, path):
    """Write a bridged graph (the artefact of step 1, main_bridged_graph.py:317-320) as a plain dict of CPU tensors.
    The reference pickles a torch_geometric Data object, which cannot be produced without PyG; ``load_graph`` reads
    both forms."""
    torch.save({"format": "bridged_gnn_b200.graph.v1", **{k: getattr(data, k).cpu() for k in data.keys()
                                                           if torch.is_tensor(getattr(data, k))}}, path)


def load_graph(path):
    """Read a bridged graph written either by ``save_graph`` (dict of tensors) or by the reference (pickled
    torch_geometric Data, ``data_bridged_graph/*.dat``)."""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        return load_pyg_dat(path)                 # a pickled PyG object needs the attribute-bag unpickler
    if isinstance(obj, dict) and "x" in obj and "edge_index" in obj:
        return Data(**{k: v for k, v in obj.items() if torch.is_tensor(v)})
    raise ValueError("%s: neither a bridged_gnn_b200 graph dict nor a pickled PyG Data" % path)
